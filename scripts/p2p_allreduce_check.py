"""Peer-memory all-reduce (csrc/k_p2p.cu) against NCCL on the same data, timing of both.  Two-shot variant: bit-exact agreement with a
rank-ordered fp32 sum.  NVLS variant (multimem.ld_reduce / multimem.st through the buffer's multicast mapping; default when the platform has
one, B4R_DISABLE_NVLS=1 selects the two-shot kernel): the switch fixes the association order, so the check is 1e-6 relative to the
rank-ordered sum plus bit-identity ACROSS ranks.
usage: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/p2p_allreduce_check.py [n_floats]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm
from bert4rec_b200 import _lib

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
lib = _lib.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1854769
dev = torch.device("cuda", torch.cuda.current_device())
buf = symm.empty(n + 8, dtype=torch.float32, device=dev)
flags = symm.empty(3 * lib.b4r_p2p_allreduce_max_world(), dtype=torch.int32, device=dev)
buf.zero_(); flags.zero_(); torch.cuda.synchronize()
hb, hf = symm.rendezvous(buf, dist.group.WORLD), symm.rendezvous(flags, dist.group.WORLD)
state = torch.zeros(8, dtype=torch.int32, device=dev)
mc = 0 if os.environ.get("B4R_DISABLE_NVLS") else int(getattr(hb, "multicast_ptr", 0) or 0)
dist.barrier()
def p2p():
    _lib.check(lib.b4r_p2p_allreduce_f32(C.c_void_p(hb.buffer_ptrs_dev), C.c_void_p(hf.buffer_ptrs_dev), C.c_void_p(mc), 0, n, rank, world,
                                         C.c_void_p(state.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
g = torch.Generator(device=dev).manual_seed(100 + rank)
for trial in range(3):
    x = torch.randn(n, device=dev, generator=g) * (1 + rank)
    allx = [torch.empty_like(x) for _ in range(world)]
    dist.all_gather(allx, x)
    want = allx[0].clone()
    for q in range(1, world):
        want += allx[q]                       # rank order, fp32: what the kernel computes
    buf[:n].copy_(x); buf[n:].fill_(7.0)
    torch.cuda.synchronize(); dist.barrier()
    p2p(); torch.cuda.synchronize()
    got = buf[:n].clone()
    if mc:
        # the switch's association order differs from rank order: bound = world * eps * sum of |terms|
        mag = allx[0].abs().clone()
        for q in range(1, world):
            mag += allx[q].abs()
        err = float(((got - want).abs() / (mag + 1e-30)).max())
        every = [torch.empty_like(got) for _ in range(world)]
        dist.all_gather(every, got)
        same = all(torch.equal(every[0], e) for e in every)
        ok = err <= world * 1.2e-7 and same
        print(f"rank {rank} trial {trial}: max |got - want| / sum|terms| = {err:.2e} (bound {world * 1.2e-7:.1e}), identical on all ranks: {same}", flush=True)
    else:
        ok = torch.equal(got, want)
    ok = ok and bool((buf[n:] == 7.0).all())
    print(f"rank {rank} trial {trial}: {'NVLS, matches and identical on all ranks' if mc else 'two-shot, bit-exact'} {ok}, err flag {int(state[7])}", flush=True)
def timeit(fn, iters=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
y = torch.randn(n, device=dev)
t_p2p = timeit(p2p)
t_nccl = timeit(lambda: dist.all_reduce(y))
if rank == 0:
    print(f"n = {n} floats ({n * 4 / 1e6:.1f} MB), world {world}: own kernel ({'NVLS one-pass' if mc else 'two-shot'}) {t_p2p:.1f} us, NCCL {t_nccl:.1f} us per all-reduce (back to back)", flush=True)
dist.destroy_process_group()
