"""Row-sharded exact kNN over NCCL: every rank builds the same synthetic database, owns rows [lo, hi), and the
sharded search + gather must be BIT-identical to the unsharded search on the full database.  Also times the sharded search.
torchrun --nproc-per-node N tools/dist_knn_check.py [rows] [queries]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "retrieval-augmented-diffusion-models_b200")]
import torch, torch.distributed as dist
from rdm_b200.knn import B200Searcher, ShardedSearcher, shard_range
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 16
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
g = torch.Generator(device=dev).manual_seed(6)
db = torch.randn((n, 512), generator=g, device=dev).to(torch.float16)
db[1234] = db[77]; db[n - 5] = db[77]                                   # duplicate rows across shards: ties must resolve to the lowest index
qids = torch.tensor([77, 3, n - 1, n // 2] + list(range(100, 100 + nq - 4)), device=dev)
lo, hi = shard_range(n, rank, world)
full = B200Searcher(db, device=dev)
shard = ShardedSearcher(B200Searcher(db[lo:hi], device=dev, idx_base=lo))
qh = torch.nn.functional.normalize(full.gather_device(qids), dim=1)
ok = True
for k in (4, 8, 20):
    i0, d0 = full.search_device(qh, k)
    i1, d1 = shard.search_device(qh, k)
    ok &= bool(torch.equal(i0, i1)) and bool(torch.equal(d0.view(torch.int32), d1.view(torch.int32)))
    ok &= bool(torch.equal(full.gather_device(i0), shard.gather_device(i1)))
for _ in range(5):
    shard.search_device(qh, 4)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
dist.barrier(); torch.cuda.synchronize(); a.record()
for _ in range(50):
    shard.search_device(qh, 4)
b.record(); torch.cuda.synchronize()
ms = torch.tensor([a.elapsed_time(b) / 50], device=dev); dist.all_reduce(ms, op=dist.ReduceOp.MAX)
flag = torch.tensor([1 if ok else 0], device=dev); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps({"world": world, "rows": n, "queries": nq, "bit_identical_to_unsharded": bool(flag.item()), "sharded_search_ms": float(ms),
                      "aggregate_gbs": n * 512 * 2 / float(ms) / 1e6}))
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
