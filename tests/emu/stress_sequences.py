"""Hand-run stress of the emulated executors (python tests/emu/stress_sequences.py, ~2 min): one handle, many calls of changing shapes --
kNN batches of 1..64 queries, CLIP batch growth / shrink, RARM context growth with mode switches and interleaved forward / sample calls,
DDIM with eta > 0 noise tables and changing batches -- every result against the oracle, every buffer behind guard zones."""
import contextlib, ctypes, os, sys, time, ast
ROOT=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, ROOT+"/retrieval-augmented-diffusion-models_b200", ROOT+"/tests", ROOT+"/tests/emu", ROOT+"/tests/golden"): sys.path.insert(0,p)
os.environ["RDM_KNN_NO_TC"]="1"
import numpy as np, torch, build_emu, ref_weights
from rdm_b200 import _lib
L=_lib.bind(ctypes.CDLL(build_emu.build()), [n for n in _lib.SIGNATURES if n.startswith(("rdm_unet_","rdm_ddim_","rdm_knn_","rdm_clip_","rdm_rarm_"))]+["rdm_last_error","rdm_launch_count"])
_lib._lib=L; _lib.resolve_device=lambda d: torch.device("cpu"); _lib.device_ctx=lambda d: contextlib.nullcontext(); _lib.stream_ptr=lambda d=None: None
from oracle import knn as oknn, clip as oclip, rarm as orarm, unet as ounet, ddim as oddim
rel=lambda a,b: float((a.double()-torch.as_tensor(b).double()).norm()/torch.as_tensor(b).double().norm())
# --- kNN: one handle, many searches of different shapes
from rdm_b200.knn import B200Searcher
rng=np.random.default_rng(5)
db=(rng.standard_normal((6000,512))*rng.uniform(0.5,8,(6000,1))).astype(np.float16)
s=B200Searcher(db, device="cpu")
for nq,k in [(1,4),(40,8),(2,24),(17,1),(64,20),(3,4)]:
    qh=oknn.normalize_queries(rng.standard_normal((nq,512)).astype(np.float32))
    idx,dist,sc=s.search_device(torch.from_numpy(qh),k,return_scores=True)
    wi,wd,ws=oknn.search(db,qh,k,return_scores=True)
    assert np.array_equal(idx.numpy(),wi) and np.array_equal(sc.numpy().view(np.int64),ws.view(np.int64)), (nq,k)
print("knn sequence ok", flush=True)
# --- CLIP: workspace growth / shrink
from rdm_b200.clip import B200Clip, cfg_from_state_dict
d=np.load(ROOT+"/tests/golden/clip_small.npz")
sd={k[3:]: torch.from_numpy(d[k]) for k in d.files if k.startswith("sd:")}
m=B200Clip("cpu", **cfg_from_state_dict(sd)); m.load_state_dict(sd); m.set_mode(0)
img=torch.from_numpy(d["image"]); tok=torch.from_numpy(d["tokens"])
for B in (1,3,2,7,1):
    ii=img[torch.arange(B)%img.shape[0]]; tt=tok[torch.arange(B)%tok.shape[0]]
    assert rel(m.encode_image(ii), oclip.encode_image(sd, ii))<1e-5 and rel(m.encode_text(tt), oclip.encode_text(sd, tt, 1))<1e-5, B
print("clip sequence ok", flush=True)
# --- RARM: context growth, interleaving forward_token and sample, mode switches
from rdm_b200.rarm import B200Rarm
g=np.load(ROOT+"/tests/golden/ref_rarm_small.npz"); cfg=ast.literal_eval(str(g["cfg_json"]))
net=B200Rarm("cpu", **cfg); rsd=ref_weights.state_dict_for(net.shapes.items(), 21); net.load_state_dict(rsd)
tg=torch.Generator().manual_seed(2)
for mode in (0,4,0):
    net.set_mode(mode)
    osd = rsd if mode==0 else ref_weights.round_dense_weights_to_fp16(rsd)
    for B,k,T in [(2,2,3),(9,5,2),(1,1,4),(4,8,3)]:
        tok=torch.randint(0,50,(B,T),generator=tg); ctx=torch.randn(B,k,128,generator=tg)
        assert rel(net.forward(tok,ctx), orarm.forward(osd,tok,ctx,cfg["n_heads"]))<3e-6, (mode,B,k,T)
        c=torch.full((B,1),49); u=torch.rand(3,B,generator=tg)
        net.set_context(ctx)
        toks=net.sample(c,3,temperature=0.8,top_k=7,guidance_scale=1.0,uniforms=u)
        want,_=orarm.sample(osd,cfg["n_heads"],c,torch.zeros((B,0),dtype=torch.long),ctx,3,0.8,7,1.0,u)
        assert torch.equal(toks[:,1:],want), (mode,B,k,T)
print("rarm sequence ok", flush=True)
# --- U-Net DDIM: eta>0 noise tables, batch changes, pred_x0
from rdm_b200.unet import B200UNet
from rdm_b200 import sampler
ref=ounet.randomize_(ounet.UNetModel(**ounet.TINY_UNET),3).eval()
un=B200UNet("cpu", **ounet.TINY_UNET); un.load_state_dict(ref.state_dict()); un.set_mode(0)
for B,H,S,eta,scale in [(1,8,2,0.5,2.0),(3,4,2,0.0,1.0),(2,8,2,1.0,2.0)]:
    x_T=torch.randn(B,4,H,H,generator=tg); c=torch.randn(B,2,512,generator=tg); uc=torch.zeros(B,2,512)
    tb=sampler.make_ddim_tables(sampler.alphas_cumprod_linear(), S, eta)
    noise=torch.randn(S,B*4*H*H,generator=tg)
    un.set_context(torch.cat([c,uc]) if scale>1 else c)
    got=un.ddim_sample(x_T,tb["timesteps"],tb["coef"],cfg_scale=scale,noise=noise if eta>0 else None)
    # oracle with the same noise
    sch=oddim.Schedule(S,eta); x=x_T.clone()
    for i,step in enumerate(np.flip(sch.timesteps)):
        ts=torch.full((B,),int(step))
        with torch.no_grad():
            if scale>1:
                o=ref(torch.cat([x]*2),torch.cat([ts]*2),torch.cat([c,uc])); e=o[B:]+scale*(o[:B]-o[B:])
            else: e=ref(x,ts,c)
        x,_=oddim.ddim_update(x,e,*sch.coeffs(S-i-1),noise=noise[i].reshape(x.shape) if eta>0 else None)
    assert rel(got,x)<1e-5,(B,H,S,eta,scale)
print("ddim sequence ok", flush=True)
