"""Stage breakdown of one create_proof (ZKB_TRACE=1 makes the session print wall-clock per stage after a stream synchronise) plus the
per-kernel-class device time (zkb_prof_*).  usage: python scripts/proof_trace.py [super|keccak] [k] [advice]"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench

kind = sys.argv[1] if len(sys.argv) > 1 else "super"
k = int(sys.argv[2]) if len(sys.argv) > 2 else 20
adv = int(sys.argv[3]) if len(sys.argv) > 3 else 128
import zkb200
ctx = zkb200.default_context(0)
pin = bench.Pinned()
sc, pk, fixed, sigma, setup = bench.build_case(kind, k, adv, pin)
prove_host, prove_dev, inst = bench.make_provers(sc, pk, pin)
for _ in range(2):
    prove_dev()
torch.cuda.synchronize()
os.environ["ZKB_TRACE"] = "1"
ctx.prof_enable(True); ctx.prof_read(0, reset=True)
t0 = time.perf_counter(); prove_dev(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(json.dumps({"shape": sc.shape, "seconds_traced": dt, "kernels": {nm: ctx.prof_read(i) for i, nm in enumerate(["ntt_tile", "msm_acc_chunk", "expr"])}}))
