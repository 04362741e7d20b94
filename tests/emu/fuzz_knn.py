"""Randomised comparison of the emulated exact searcher (csrc/knn.cu, SIMT path) with the oracle: random row widths, sizes, query counts, k,
blocks of exact duplicates and near-duplicate clusters.  Not collected by pytest (run by hand: python tests/emu/fuzz_knn.py); 40 cases, ~1 min."""
import contextlib, ctypes, os, sys, time
ROOT=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, ROOT+"/retrieval-augmented-diffusion-models_b200", ROOT+"/tests", ROOT+"/tests/emu"): sys.path.insert(0,p)
os.environ["RDM_KNN_NO_TC"]="1"
import numpy as np, torch, build_emu
from rdm_b200 import _lib
L=_lib.bind(ctypes.CDLL(build_emu.build()), [n for n in _lib.SIGNATURES if n.startswith("rdm_knn_")]+["rdm_last_error","rdm_launch_count"])
_lib._lib=L; _lib.resolve_device=lambda d: torch.device("cpu"); _lib.device_ctx=lambda d: contextlib.nullcontext(); _lib.stream_ptr=lambda d=None: None
from rdm_b200.knn import B200Searcher
from oracle import knn as oknn
rng=np.random.default_rng(123)
bad=0
for case in range(40):
    dtype=[np.float16,np.float32][case%2]
    d=int(rng.choice([256,512,768,1024]))
    n=int(rng.integers(24,4000)); nq=int(rng.integers(1,21)); k=int(rng.integers(1,25))
    db=(rng.standard_normal((n,d))*rng.uniform(0.5,8,(n,1))).astype(dtype)
    kind=case%4
    if kind==1 and n>300:   # block of exact duplicates
        m=int(rng.integers(2,min(n//2,1500))); s0=int(rng.integers(0,n-m)); db[s0:s0+m]=db[s0]
    if kind==2 and n>500:   # near-duplicate cluster
        s0=int(rng.integers(0,n-400)); db[s0:s0+400]=db[s0]+(rng.standard_normal((400,d))*1e-3).astype(dtype)
    q=rng.standard_normal((nq,d)).astype(np.float32)
    q[0]=db[int(rng.integers(0,n))].astype(np.float32)
    if kind==1 and n>300: q[nq-1]=db[s0].astype(np.float32)
    if kind==2 and n>500: q[nq-1]=db[s0].astype(np.float32)
    qh=oknn.normalize_queries(q)
    s=B200Searcher(db, device="cpu")
    idx,dist,sc=s.search_device(torch.from_numpy(qh),k,return_scores=True)
    wi,wd,ws=oknn.search(db,qh,k,return_scores=True)
    ok=np.array_equal(idx.numpy(),wi) and np.array_equal(sc.numpy().view(np.int64),ws.view(np.int64))
    if not ok:
        bad+=1; print("MISMATCH case",case,dtype.__name__,"d",d,"n",n,"nq",nq,"k",k,"kind",kind, flush=True)
print("fuzz done, mismatches:",bad)
