#!/usr/bin/env python
"""bench.py - ClipLoss fwd+bwd samples/s at global batch 32768, d=512, bf16 (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl clipk|reference]

One process per GPU (under torchrun for N > 1; RANK / LOCAL_RANK / WORLD_SIZE from the environment).  The global
batch is fixed, so the N-GPU runs are a strong-scaling series: each rank owns 32768 / N rows.

A step is one ClipLoss(local_loss=True, gather_with_grad=True) forward + backward on synthetic unit-norm embeddings
(positives at cos ~0.3).  Every step is bracketed by its own pair of CUDA events on the
launching stream; between steps a 256 MiB buffer is overwritten to flush the 126 MB L2 (outside the event pairs).
`value` has the inputs resident in HBM; `e2e` runs the same call from pinned host buffers (H2D of both feature
matrices and D2H of the loss inside the timed region), with the copy of the next step's inputs prefetched on a copy
stream while the current step runs (`e2e.serial_*` is the same without the overlap).  rank 0 prints ONE JSON line.

--impl reference times the reference's own CPU arithmetic (torch CPU matmul + cross_entropy + autograd, restated in
oracle/cliploss_oracle.py::TorchPort because /root/reference does not exist on the GPU box) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "megatron-clip_b200"))

GLOBAL_BATCH = 32768
DIM = 512
LOGIT_SCALE = 1.0 / 0.07
METRIC = "ClipLoss fwd+bwd samples/s @ global batch 32K, d=512"
CPU_SAMPLE_BATCH = 8192   # bounded CPU sample: fwd+bwd cost grows with batch^2


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            j = json.load(f)
        return {"tflops_sustained": j.get("bf16_tflops_sustained", 1401.9), "tflops_burst": j.get("bf16_tflops", 1661.2),
                "hbm_gbs": j.get("hbm_gbs", 6542.7), "source": "measured (MEASURED_PEAKS.json)"}
    return {"tflops_sustained": 1400.0, "tflops_burst": 1590.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi polling the clocks of one GPU during the timed region.  The samples go to a file that is read after
    the run: a reader thread in this process would fight the launching thread for the interpreter lock."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path, self.fh = index, None, None, None

    def start(self):
        import tempfile
        try:
            self.fh = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.path = self.fh.name
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", os.environ.get("CLIPK_BENCH_SMI_MS", "50")],
                                         stdout=self.fh, stderr=subprocess.DEVNULL, text=True)
            time.sleep(0.3)         # let the first samples land before the timed region starts
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.fh.close()
        with open(self.path) as f:
            rows = [[c.strip() for c in line.split(",")] for line in f if line.strip()]
        os.unlink(self.path)
        sm, mx, power, reasons = [], None, [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
                power.append(float(r[3]))
                for n, v in zip(names, r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except (ValueError, IndexError):
                pass
        # clocks under load = the samples taken while the GPU drew the most power (busier half)
        order = sorted(range(len(sm)), key=lambda i: power[i])
        load = sorted(sm[i] for i in order[len(order) // 2:])
        med = load[len(load) // 2] if load else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": max(power) if power else None}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_port_run(steps, warmup, batch=CPU_SAMPLE_BATCH):
    """Reference arithmetic on host cores: returns (seconds per sample-step, cores)."""
    import numpy as np
    import torch
    from oracle import cliploss_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    x, t = O.synthetic_features(batch, DIM, seed=1234)
    I, T, s = torch.from_numpy(x), torch.from_numpy(t), torch.tensor(LOGIT_SCALE)
    port = O.TorchPort()
    for _ in range(warmup):
        port.fwd_bwd(I, T, s)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        port.fwd_bwd(I, T, s)
        ts.append(time.perf_counter() - t0)
    return float(np.mean(ts)), cores


def cpu_baseline_obj(sec_per_step, cores, batch):
    # cost per step grows with batch^2 (two batch x batch logit matrices); scale the sample to the 32K workload
    full = sec_per_step * (GLOBAL_BATCH / batch) ** 2
    return {"value": GLOBAL_BATCH / full, "unit": "samples/s", "cores": cores, "kind": "port",
            "sample": (f"torch-CPU fp32 ClipLoss fwd+bwd at batch {batch}, d={DIM} ({sec_per_step * 1e3:.0f} ms/step), "
                       f"scaled by (32768/{batch})^2 to the global-batch-32768 step")}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 5))
    sec, cores = cpu_port_run(steps, min(args.warmup, 1))
    cb = cpu_baseline_obj(sec, cores, CPU_SAMPLE_BATCH)
    full_ms = GLOBAL_BATCH / cb["value"] * 1e3
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "samples/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": full_ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "ClipLoss fwd+bwd, global batch 32768, d=512 (BASELINE.json configs[1]), CPU arithmetic"},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------ GPU arm
def run_clipk(args):
    import torch
    import torch.distributed as dist
    from clipk import ClipLoss, ops
    from oracle import cliploss_oracle as O   # synthetic input generator + cpu_baseline leg only

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert GLOBAL_BATCH % world == 0
    b = GLOBAL_BATCH // world

    x, t = O.synthetic_features(b, DIM, seed=1234, rank=rank)
    I_host = torch.from_numpy(x).bfloat16().pin_memory()
    T_host = torch.from_numpy(t).bfloat16().pin_memory()
    I = I_host.to(dev).requires_grad_(True)
    T = T_host.to(dev).requires_grad_(True)
    S = torch.tensor(LOGIT_SCALE, device=dev, requires_grad=True)
    loss_mod = ClipLoss(local_loss=True, gather_with_grad=True, cache_labels=True, rank=rank, world_size=world)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()

    def step_resident():
        I.grad = T.grad = S.grad = None
        loss = loss_mod(I, T, S)
        loss.backward()
        return loss

    def step_e2e():
        i = I_host.to(dev, non_blocking=True).requires_grad_(True)
        tt = T_host.to(dev, non_blocking=True).requires_grad_(True)
        S.grad = None
        loss = loss_mod(i, tt, S)
        loss.backward()
        loss_host.copy_(loss.detach(), non_blocking=True)
        return loss

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        # every step has its own event pair; the L2 flush between steps is enqueued outside the pairs and the host
        # does not synchronise inside the timed region, so launch latency hides behind the previous step's kernels
        launches = 0
        pairs = []
        for _ in range(steps):
            flush.fill_(1)                          # L2 flush, outside the event pair
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            la = ops.gpu_launches()
            e0.record()
            fn()
            e1.record()
            launches += ops.gpu_launches() - la
            pairs.append((e0, e1))
        torch.cuda.synchronize()
        total_ms = sum(e0.elapsed_time(e1) for e0, e1 in pairs)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        tt = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)   # max over ranks
        return tt.item() / steps, launches

    def timed_pipelined(steps, warmup):
        """e2e with the input copies double-buffered, the way a host-fed training loop prefetches: the H2D copy of
        step k+1's features runs on a copy stream while step k's kernels run (it may not start before step k's timed
        region has begun), and step k+1 waits for it inside its own event pair.  Same per-step event pairs, L2 flush
        and max over ranks as `timed`; the D2H read of the loss stays inside each pair."""
        main = torch.cuda.current_stream(dev)
        copy_stream = torch.cuda.Stream(device=dev)
        bufs = [(torch.empty(I_host.shape, dtype=I_host.dtype, device=dev), torch.empty(T_host.shape, dtype=T_host.dtype, device=dev))
                for _ in range(2)]
        ready = [torch.cuda.Event(), torch.cuda.Event()]

        def prefetch(k, after):
            copy_stream.wait_event(after)
            with torch.cuda.stream(copy_stream):
                bufs[k % 2][0].copy_(I_host, non_blocking=True)
                bufs[k % 2][1].copy_(T_host, non_blocking=True)
                ready[k % 2].record(copy_stream)

        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        total = warmup + steps
        begin = torch.cuda.Event()
        begin.record(main)
        prefetch(0, begin)
        pairs = []
        for k in range(total):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(main)
            if k + 1 < total:
                prefetch(k + 1, e0)      # buffer (k+1)%2 was last read by step k-1, which precedes e0 on `main`
            main.wait_event(ready[k % 2])
            i = bufs[k % 2][0].detach().requires_grad_(True)
            tt = bufs[k % 2][1].detach().requires_grad_(True)
            S.grad = None
            loss = loss_mod(i, tt, S)
            loss.backward()
            loss_host.copy_(loss.detach(), non_blocking=True)
            e1.record(main)
            if k >= warmup:
                pairs.append((e0, e1))
        torch.cuda.synchronize()
        total_ms = sum(e0.elapsed_time(e1) for e0, e1 in pairs)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        tt = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return tt.item() / steps, float(loss_host)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    ms, launches = timed(step_resident, args.steps, args.warmup)
    loss_resident = float(step_resident().detach())
    e2e_steps = max(3, args.steps // 2)
    ms_e2e_serial, _ = timed(step_e2e, e2e_steps, 3)
    # the pipelined figure is the headline only if it ran and reproduced the loss of the resident inputs (every rank
    # must agree on that, the collectives inside a step need all of them)
    ms_e2e, e2e_mode = ms_e2e_serial, "serial: H2D, step and D2H on one stream"
    if os.environ.get("CLIPK_BENCH_E2E", "pipelined") == "pipelined":
        ms_pipe, loss_pipe = timed_pipelined(e2e_steps, 3)
        ok = torch.tensor([1 if abs(loss_pipe - loss_resident) <= 1e-5 * abs(loss_resident) else 0], device=dev)
        if world > 1:
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 1:
            ms_e2e = ms_pipe
            e2e_mode = ("double-buffered: the H2D copy of step k+1 runs on a copy stream during step k (not before "
                        "step k's timed region begins); step k+1 waits for it inside its own event pair")
        else:
            e2e_mode += f"; pipelined run rejected (loss {loss_pipe} != {loss_resident})"

    # ---- per-kernel breakdown of one rank's step (events around each C-ABI call), for the roofline
    be = ops._backend()
    N = GLOBAL_BATCH
    breakdown = {}
    with torch.no_grad():
        sc = S.detach().reshape(1).float()
        if world > 1:
            t_all = torch.empty(N, DIM, dtype=torch.bfloat16, device=dev)
            dist.all_gather_into_tensor(t_all, T.detach())
        else:
            t_all = T.detach()
        X, Y = be.prepare(I.detach()), be.prepare(t_all)
        off = rank * b if world > 1 else 0

        def ev(fn, reps=5):
            out = None
            for _ in range(2):      # warm up; the outputs are dropped before the next call so that the caching
                out = None          # allocator reuses their memory (a cudaMalloc inside the event pair would be timed)
                out = fn()
            torch.cuda.synchronize()
            tot = 0.0
            for _ in range(reps):
                out = None
                flush.fill_(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                out = fn()
                e1.record()
                torch.cuda.synchronize()
                tot += e0.elapsed_time(e1)
            return tot / reps, out

        parts = torch.empty(1, 3, N, dtype=torch.float32, device=dev)
        breakdown["fwd_both_ms"], (rstats, pos, _) = ev(lambda: be.fwd_both(X, Y, sc, off, col_out=parts[0]))
        gparts = parts
        if world > 1:
            gparts = torch.empty(world, 3, N, dtype=torch.float32, device=dev)
            dist.all_gather_into_tensor(gparts, parts)
        lse_row, lse_col, sums = be.finalize(rstats, pos, gparts, off)
        breakdown["to_f16_ms"], (Xg, Yg) = ev(lambda: (be.prepare_grad(X), be.prepare_grad(Y)))
        gscale = torch.tensor([1.0 / (2 * b)], device=dev)
        breakdown["bwd_ms"], (dXa, dYa) = ev(lambda: be.bwd(X, Y, Xg, Yg, sc, off, lse_row, lse_col, 1.0, 1.0, gscale, True, True))
        if world > 1:
            # the three collectives of a step, timed alone (they are issued on the same stream as the kernels)
            tl = T.detach()
            breakdown["all_gather_T_ms"], _ = ev(lambda: ops._all_gather_rows(tl, world))
            breakdown["all_gather_colstats_ms"], _ = ev(lambda: ops._all_gather_rows(parts, world))
            breakdown["reduce_scatter_dT_ms"], _ = ev(lambda: ops._reduce_scatter_rows(dYa, world))
    # nvidia-smi cannot sample faster than every ~50 ms and the timed region lasts ~0.1 s, so the sampler stays on from
    # before the timed region until the end of a short untimed continuation of the same step (all ranks take part: the
    # step holds collectives); `clocks` is the busier half of those samples.
    probe_s = float(os.environ.get("CLIPK_BENCH_CLOCK_PROBE_S", "0.8"))
    probe_steps = 0
    try:
        n_probe = int(min(400, probe_s / max(ms * 1e-3, 1e-4)))
        for _ in range(n_probe):
            step_resident()
            probe_steps += 1
        torch.cuda.synchronize()
    except Exception as e:      # the probe must never cost the run its result line
        print(f"clock probe stopped: {e!r}", file=sys.stderr)
    clocks = sampler.stop() if sampler else None
    if clocks is not None:
        clocks["window"] = (f"timed region, e2e legs and per-kernel breakdown, then {probe_steps} more untimed steps "
                            f"(~{probe_s} s) of the same workload")
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    f_alg = 6.0 * b * N * DIM                                  # three dense passes over the b x N block
    f_exec = 8.0 * b * N * DIM                                 # issued: 1 fwd sweep (rows + columns) + recompute + 2 gradient GEMMs
    achieved = f_alg / (ms * 1e-3) / 1e12
    gemm_ms = breakdown["fwd_both_ms"] + breakdown["bwd_ms"]
    # per-kernel view, timed live above with CUDA events around each C entry (the small kernels of an entry included)
    unit = 2.0 * b * N * DIM                                   # one dense pass over the block
    kernels = [
        {"entry": "clipk_fwd_both (norm2_max + fwd_sweep_kernel x2 + fwd_merge)", "algorithmic_flops": unit,
         "ms": breakdown["fwd_both_ms"], "tflops": unit / (breakdown["fwd_both_ms"] * 1e-3) / 1e12},
        {"entry": "clipk_bwd (grad_sweep_kernel + gemm_pair_kernel per panel)", "algorithmic_flops": 2 * unit,
         "executed_mma_flops": 3 * unit, "ms": breakdown["bwd_ms"],
         "tflops": 2 * unit / (breakdown["bwd_ms"] * 1e-3) / 1e12,
         "executed_tflops": 3 * unit / (breakdown["bwd_ms"] * 1e-3) / 1e12},
    ]
    for k in kernels:
        k["frac_of_peak"] = k.get("executed_tflops", k["tflops"]) / pk["tflops_sustained"]
    # DRAM traffic per step from the committed ncu --set full captures (profiles/README.md, N = 32768 on one GPU):
    # forward sweep 85 + 46 MB; per panel (5632 x 16384), recompute 29 + 135 MB and gradient GEMMs 339 + 63 MB; 12 panels
    traffic = (131e6 + 12 * (164e6 + 402e6)) if (world == 1) else None
    roofline = {
        "bound": "tensor", "achieved": achieved, "peak": pk["tflops_sustained"], "unit": "TFLOP/s",
        "frac": achieved / pk["tflops_sustained"], "traffic": traffic,
        "traffic_note": "DRAM bytes per step (read + write) summed over the tensor-core kernels, ncu --set full, "
                        "profiles/r01m_*; algorithmic operand bytes are 64 MB - the rest is the fp16 "
                        "softmax-gradient panel streaming through HBM (6.7 GB) and 46 MB of column partials",
        "kernel": "whole step (SURVEY 8d: F_alg = 6 b N d over t_step); dominant kernels: clipk::gemm_pair_kernel "
                  "(gradient GEMMs, 49 % of the step), grad_sweep_kernel (26 %), fwd_sweep_kernel (21 %); tcgen05 "
                  "cta_group::2 256x256x64 tiles",
        "algorithmic_flops_per_step_per_gpu": f_alg, "executed_mma_flops_per_step_per_gpu": f_exec,
        "executed_tflops_in_gemm_kernels": f_exec / (gemm_ms * 1e-3) / 1e12,
        "gemm_kernels_share_of_step": gemm_ms / ms, "breakdown_ms": breakdown, "kernels": kernels,
        "peak_source": pk["source"], "peak_burst": pk["tflops_burst"],
    }

    cb = None
    if world == 1 and not os.environ.get("CLIPK_BENCH_QUICK"):     # development runs may skip the CPU leg
        sec, cores = cpu_port_run(2, 1)
        cb = cpu_baseline_obj(sec, cores, CPU_SAMPLE_BATCH)

    out = {
        "metric": METRIC, "value": GLOBAL_BATCH / (ms * 1e-3), "unit": "samples/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "ClipLoss local_loss=True gather_with_grad=True fwd+bwd, ViT-B/32 embeddings "
                               "(BASELINE.json configs[1])", "global_batch": GLOBAL_BATCH, "local_batch": b, "d": DIM,
                   "logit_scale": LOGIT_SCALE, "parallelism": f"dp{world}",
                   "l2": "256 MiB buffer overwritten between timed steps (L2 flush), outside the per-step event pairs"},
        "clocks": clocks,
        "e2e": {"value": GLOBAL_BATCH / (ms_e2e * 1e-3), "unit": "samples/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": 2 * b * DIM * 2, "d2h_bytes_per_step": 4, "mode": e2e_mode,
                "serial_ms_per_step": ms_e2e_serial, "serial_value": GLOBAL_BATCH / (ms_e2e_serial * 1e-3)},
        "gpu_launches": launches,
        "roofline": roofline,
        "cpu_baseline": cb,
    }
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    # exactly ONE JSON line may reach stdout: libraries (NCCL's version banner) write there too, so everything but the
    # final line is sent to stderr at the file-descriptor level
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="clipk", choices=["clipk", "reference"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "clipk" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_clipk(args)


if __name__ == "__main__":
    main()
