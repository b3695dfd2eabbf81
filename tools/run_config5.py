"""BASELINE config 5: 1 Mi trajectories (N = 20, Ts = 0.01, T = 1200) block-partitioned over the GPUs of one box.
Launch under torchrun.  Rows stay in HBM as per-GPU shards (fp64: 134 KB per trajectory -> 17.6 GB per GPU at 8 GPUs);
rank 0 additionally writes the CSV of the first 5000 ids.  Time = max over ranks (CUDA events)."""
import os, sys, time
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, ".")
import bench
import trajectory_generation_b200 as tg
from trajectory_generation_b200 import _lib, distributed as tgd

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
if world > 1:
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", lr))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
TOTAL = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
T = 1200
lo, hi = tgd.shard_range(TOTAL, rank, world)
B = hi - lo
# scenarios, initial states and steady-state inputs: generated on the device, resident (tg_make_scenarios), the spline /
# sinusoid / parabola mix of configs 2 and 3 by id mod 3
gen = tg.ClosedLoopGenerator(device=lr, N=20, Ts=0.01, plant=tg.PLANT_GEN1, vref_advance=True)
stream = torch.cuda.Stream(device=dev); gen.set_stream(stream.cuda_stream)
rules = tg.scenario_rules(cycle=(tg.PATH_SPLINE, tg.PATH_SINE, tg.PATH_PARABOLA))
P = rules.spl_knots - 1
d_x0 = torch.empty((B, 6), dtype=torch.float64, device=dev); d_u0 = torch.empty((B, 2), dtype=torch.float64, device=dev)
d_spec = torch.empty(B * _lib.REF_SPEC_DTYPE.itemsize, dtype=torch.uint8, device=dev)
d_brk = torch.empty(B * P, dtype=torch.float64, device=dev); d_coef = torch.empty((B * P, 4), dtype=torch.float64, device=dev)
clean = torch.empty((B, T + 1, 6), dtype=torch.float64, device=dev); noisy = torch.empty_like(clean)
U = torch.empty((B, T, 2), dtype=torch.float64, device=dev)
scnt = torch.zeros((B, 6), dtype=torch.int32, device=dev); its = torch.zeros(B, dtype=torch.int64, device=dev)
L = _lib.load()
import ctypes
t_s = time.perf_counter()
with torch.cuda.stream(stream):
    _lib.check(L.tg_make_scenarios(gen.handle, B, lo, ctypes.byref(rules), d_x0.data_ptr(), d_u0.data_ptr(), d_spec.data_ptr(), d_brk.data_ptr(), d_coef.data_ptr()))
torch.cuda.synchronize(dev)
t_scen = time.perf_counter() - t_s
def launch(nb, t):
    _lib.check(L.tg_closed_loop(gen.handle, nb, t, d_x0.data_ptr(), d_u0.data_ptr(), d_spec.data_ptr(), d_brk.data_ptr(), d_coef.data_ptr(), lo,
                                clean.data_ptr(), noisy.data_ptr(), U.data_ptr(), scnt.data_ptr(), its.data_ptr()))
with torch.cuda.stream(stream):
    launch(min(B, 2048), 20)          # warm-up
torch.cuda.synchronize(dev)
if world > 1: dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(stream):
    e0.record(stream); launch(B, T); e1.record(stream)
torch.cuda.synchronize(dev)
ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
agg = torch.cat([scnt.sum(0).to(torch.int64), its.sum()[None]])
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX); dist.all_reduce(agg)
if rank == 0:
    steps = TOTAL * T
    print(f"config 5  {TOTAL} trajectories x {T} steps on {world} GPU(s): {ms.item()/1e3:.2f} s (max over ranks) = {steps/(ms.item()*1e-3):.3e} MPC steps/s; "
          f"shard {B} trajectories = {(2*clean.numel()+U.numel())*8/1e9:.1f} GB of rows per GPU in HBM; statuses {dict(zip(tg.STATUS_STRINGS, agg[:6].tolist()))}; "
          f"mean ADMM iterations/step {agg[6].item()/steps:.2f}; scenario generation on the device {t_scen*1e3:.1f} ms per rank")
    n5 = min(5000, B)
    res = {"clean": clean[:n5].cpu().numpy(), "noisy": noisy[:n5].cpu().numpy(), "U": U[:n5].cpu().numpy()}
    t = time.perf_counter(); tg.write_csv(res, 0.01, "/tmp/c5_clean.csv", "/tmp/c5_noisy.csv"); print(f"          CSV of the first {n5} ids: {time.perf_counter()-t:.1f} s")
if world > 1:
    dist.barrier(); dist.destroy_process_group()
