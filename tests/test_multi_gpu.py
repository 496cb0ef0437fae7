"""2-GPU sharded sweeps vs the same chain on one GPU (needs >= 2 CUDA devices; run with
`gpurun --gpus 2`).  Per-function draws are keyed by the GLOBAL function index and the globals by
(seed, iteration), so the sharded chain follows the single-GPU chain up to summation order."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _setup(rank, world, n_total=4096):
    import bayesfmmm_b200 as bf
    from bayesfmmm_b200.engine import FUNCTIONAL
    from oracle import oracle as orc
    from tests import synth
    s = synth.functional_common(seed=21, n=n_total, T=64, K=3, P=10, M=2, sigma_sq=0.04)
    lo, hi = (n_total * rank) // world, (n_total * (rank + 1)) // world
    eng = bf.Engine(model=FUNCTIONAL, n=hi - lo, K=3, P=10, M=2, y=s["y"][lo:hi], B=s["B"], T=64,
                    device=rank, global_offset=lo)
    eng.set_state(s["Z"][lo:hi], s["chi"][lo:hi])
    smp = bf.Sampler(eng, hyper=bf.default_hyper(True, a_Z_PM=2000.0), n_total=n_total, Pmat=orc.pmat_rw1(10), seed=7)
    par = s["par"]
    smp.set(nu=par["nu"], Phi=par["Phi"], sigma_sq=0.04, pi=s["pi"], alpha3=1.0)
    return bf, eng, smp, lo, hi


def _summary(smp):
    g = smp.get()
    return np.concatenate([g["nu"].ravel(), g["Phi"].ravel(), g["pi"], [g["alpha3"], g["sigma_sq"], g["loglik"]],
                           g["tau"], g["delta"].ravel()])


def _worker(rank, world, port, q, n_sweeps, hook="python"):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    if hook.endswith("_dev"):       # device-resident sweep: the exchanges are queued on the stream, no host read-back
        os.environ["BFMMM_DEVICE_GLOBALS"] = "1"
        hook = hook[:-4]
    if hook.endswith("_sep"):       # peer-memory exchange as a kernel of its own (not fused into the statistics pass's epilogue)
        os.environ["BFMMM_NO_FUSED_EXCHANGE"] = "1"
        hook = hook[:-4]
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    bf, eng, smp, lo, hi = _setup(rank, world)
    ext = torch.cuda.ExternalStream(eng.stream, device=torch.device("cuda", rank))

    class _Buf:
        def __init__(self, ptr, ln):
            self.__cuda_array_interface__ = {"shape": (ln,), "typestr": "<f8", "data": (ptr, False), "version": 3}
    ptr, ln = eng.stats_buffer()
    t = torch.as_tensor(_Buf(ptr, ln), device=torch.device("cuda", rank))

    def allreduce(p, l, stream):
        off = (p - ptr) // 8
        with torch.cuda.stream(ext):
            dist.all_reduce(t[off:off + l])
    dev = torch.device("cuda", rank)
    if hook == "python":            # generic hook through torch.distributed
        smp.set_allreduce(allreduce)
    elif hook == "nccl":            # native ncclAllReduce (csrc/nccl_hook.cu)
        def exchange_id(idb):
            tt = torch.zeros(128, dtype=torch.uint8, device=dev)
            if rank == 0:
                tt.copy_(torch.frombuffer(bytearray(idb), dtype=torch.uint8))
            dist.broadcast(tt, 0)
            return bytes(tt.cpu().numpy().tobytes())
        smp.enable_nccl(rank, world, exchange_id)
    else:                           # one-shot all-reduce over NVLink peer memory (csrc/p2p_hook.cu)
        def allgather(h):
            mine = torch.frombuffer(bytearray(h), dtype=torch.uint8).to(dev)
            out = [torch.zeros(64, dtype=torch.uint8, device=dev) for _ in range(world)]
            dist.all_gather(out, mine)
            return b"".join(bytes(o.cpu().numpy().tobytes()) for o in out)
        smp.enable_p2p(rank, world, ln, allgather)
        dist.barrier()
    for _ in range(n_sweeps):
        smp.step(bf.SWEEP_FULL)
    Z, chi = eng.get_state()
    q.put((rank, _summary(smp), Z, chi, smp.last_accept))
    dist.barrier()
    torch.cuda.synchronize()
    smp.close(); eng.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("hook", ["python", "nccl", "p2p", "p2p_sep", "nccl_dev", "p2p_dev"])
def test_two_gpu_chain_follows_single_gpu_chain(hook):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    n_sweeps = 3
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, n_sweeps, hook)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(2)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(res[0][1], res[1][1])           # identical globals on both ranks
    bf, eng, smp, lo, hi = _setup(0, 1)
    for _ in range(n_sweeps):
        smp.step(bf.SWEEP_FULL)
    single = _summary(smp)
    Z1, chi1 = eng.get_state()
    assert np.max(np.abs(res[0][1] - single) / (1 + np.abs(single))) < 1e-8
    Z2 = np.concatenate([res[0][2], res[1][2]]); chi2 = np.concatenate([res[0][3], res[1][3]])
    same = np.all(np.abs(Z2 - Z1) < 1e-9, axis=1)
    assert same.mean() > 0.999                             # a knife-edge accept may flip a row
    assert np.max(np.abs(chi2 - chi1)[same]) < 1e-7
    assert res[0][4] == res[1][4] and abs(res[0][4] - smp.last_accept) <= 2
    smp.close(); eng.close()
