"""bench.py -- the contract benchmark (see DESIGN.md "Measurement").

Workload (BASELINE.json configs[1]): GPTQ INT4 g128 calibration of a random-init Llama-2-7B (fp16) on 128
synthetic sequences of 2048 tokens.  GPTQ is block-sequential, so a STEP is one decoder block of the real
32-block model processed exactly as the algorithm does it: forward #1 with Hessian hooks over all 262 144
calibration tokens, Cholesky inverse factor, column loop + lazy updates for the 7 linears, forward #2 with the
quantised weights (its outputs are the next block's inputs), packing.  `--steps 29 --warmup 3` is the whole model.

    value  = calibration tokens / second of the WHOLE 32-block model = 262144 / (ms_per_step * 32 / 1000)
             (inputs resident in HBM when the timed region starts)
    e2e    = the same, through the public prepare()/convert() engine with the block weights in pinned HOST
             memory: every step copies its block H2D and its packed result D2H inside the timed region
    decode = Llama-2-7B INT4 decode tokens/s through the 224 packed linears (WeightOnlyLinear.forward, batch 1)
    roofline         = the dominant calibration kernel (Hessian SYRK, tensor bound)
    roofline_decode  = the dequant-GEMV (HBM bound)
    cpu_baseline     = the oracle port (reference arithmetic on torch CPU ops) on a bounded sample, host cores

`--impl reference` times the reference's CPU arithmetic (oracle port; the Python reference cannot travel to the
GPU box) on the same config with all host threads, rank 0 only.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HIDDEN, INTER, LAYERS, HEADS, VOCAB = 4096, 11008, 32, 32, 32000
N_SAMPLES, SEQ = 128, 2048
TOKENS = N_SAMPLES * SEQ
LINEARS = [("q", HIDDEN, HIDDEN), ("k", HIDDEN, HIDDEN), ("v", HIDDEN, HIDDEN), ("o", HIDDEN, HIDDEN),
           ("gate", INTER, HIDDEN), ("up", INTER, HIDDEN), ("down", HIDDEN, INTER)]


_T0 = time.perf_counter()


def log(msg):
    if os.environ.get("RANK", "0") == "0":
        print(f"[bench {time.perf_counter() - _T0:7.1f}s] {msg}", file=sys.stderr, flush=True)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        d["source"] = "measured (MEASURED_PEAKS.json)"
        return d
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        sm = sorted(int(float(r[1])) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit())
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        mx = int(float(self.rows[0][2])) if self.rows else None
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))


# ------------------------------------------------------------------------------------------------ model
def build_llama(dev, n_layers=LAYERS):
    from transformers import LlamaConfig, LlamaForCausalLM

    cfg = LlamaConfig(hidden_size=HIDDEN, intermediate_size=INTER, num_hidden_layers=n_layers,
                      num_attention_heads=HEADS, num_key_value_heads=HEADS, vocab_size=VOCAB,
                      max_position_embeddings=4096, tie_word_embeddings=False)
    torch.manual_seed(0)
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float16)
    try:
        with torch.device(dev):
            model = LlamaForCausalLM(cfg)
    finally:
        torch.set_default_dtype(old)
    model.eval()
    model.config.use_cache = False
    return model


def calib_ids(dev, lo, hi):
    g = torch.Generator().manual_seed(1234)
    ids = [torch.randint(0, VOCAB, (1, SEQ), generator=g) for _ in range(N_SAMPLES)]
    return [x.to(dev) for x in ids[lo:hi]]


# ------------------------------------------------------------------------------------------------ our arm
def run_b200(args):
    import torch.distributed as dist

    import neural_compressor_b200.quantization as Q
    from neural_compressor_b200 import _lib, ops

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    pk = peaks()
    W, K = args.warmup, args.steps
    n_blocks_needed = min(LAYERS, (2 * (W + K) if args.e2e else (W + K)) + (1 if args.phases else 0))
    log(f"building Llama-2-7B-shape model with {n_blocks_needed} of {LAYERS} blocks on {dev}")
    model = build_llama(dev, n_blocks_needed)
    log("model built")

    cfg = Q.GPTQConfig(bits=4, group_size=128, use_sym=True, act_order=False, block_size=128, percdamp=0.01)
    model = Q.prepare(model, cfg)
    # strong scaling: the 128 calibration sequences are split over the ranks (Hessians all-reduced)
    from neural_compressor_b200.algorithms.gptq import shard_range

    lo, hi = shard_range(N_SAMPLES, rank, world)
    with torch.no_grad():
        for ids in calib_ids(dev, lo, hi):
            model(ids)
    log(f"captured block-0 inputs of {hi - lo} sequences")
    engine = model.quantizer.gptq_quantizer
    engine.remove_prepare_for_calibration()
    engine.world_size, engine.rank = world, rank
    hess_events = []
    ops.PROFILE_HOOK = hess_events  # (start, end, flops) per Hessian launch

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_blocks(first, count, host_resident):
        if host_resident:
            for b in range(first, first + count):
                blk = engine.blocks_info["transformers"][b].to("cpu")
                for p_ in blk.parameters():
                    p_.data = p_.data.pin_memory()
        engine.offload_packed_to_host = host_resident
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = lib.b200woq_launch_count()
        s.record()
        torch.cuda.nvtx.range_push("timed_steps")  # lets `ncu --nvtx --nvtx-include "timed_steps/"` list exactly these
        with torch.no_grad():
            for b in range(first, first + count):
                engine.quantize_block(b)
        torch.cuda.nvtx.range_pop()
        e.record()
        barrier()
        ms = s.elapsed_time(e)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms, lib.b200woq_launch_count() - l0

    with torch.no_grad():
        for b in range(W):  # warm-up steps (blocks 0..W-1)
            engine.quantize_block(b)
            torch.cuda.synchronize()
            log(f"warm-up block {b} done")
    hess_events.clear()
    sampler = ClockSampler(local)
    sampler.start()
    ms_total, launches = timed_blocks(W, K, host_resident=False)
    clocks = sampler.stop()
    ms_step = ms_total / K
    log(f"timed {K} blocks: {ms_step:.1f} ms/step")
    value = TOKENS / (ms_step * LAYERS / 1000.0)

    # dominant kernel: Hessian SYRK
    torch.cuda.synchronize()
    h_ms = [s_.elapsed_time(e_) for (s_, e_, _) in hess_events]
    h_fl = [f for (_, _, f) in hess_events]
    roofline = None
    if h_ms:
        ach = sum(h_fl) / (sum(h_ms) / 1e3) / 1e12
        peak = pk["bf16_tflops_sustained"]
        roofline = dict(bound="tensor", kernel="hessian_syrk_tc_kernel (tcgen05 + TMA + TMEM)", achieved=round(ach, 1), peak=peak,
                        unit="TFLOP/s", frac=round(ach / peak, 4), traffic=syrk_traffic_per_launch(h_fl),
                        traffic_source="dram__bytes_read.sum + dram__bytes_write.sum per launch, `ncu --set full` capture "
                                       "summarised in profiles/ (launch-weighted mean over the C=4096 and C=11008 launches)",
                        peak_source=pk["source"] + ", sustained",
                        launches=len(h_ms), avg_launch_ms=round(sum(h_ms) / len(h_ms), 4),
                        share_of_step=round(sum(h_ms) / ms_total, 3),
                        algorithmic="T*C*(C+128) flops per launch: the symmetric half of the reference's 2*T*C^2")
    ops.PROFILE_HOOK = None

    phases = None
    if args.phases and W + K < n_blocks_needed:
        engine.profile = True
        for k_ in engine.timing:
            engine.timing[k_] = 0.0
        with torch.no_grad():
            engine.quantize_block(W + K)
        engine.profile = False
        phases = {k_: round(v_ * 1e3, 1) for k_, v_ in engine.timing.items()}
        log(f"phase breakdown of one extra block (ms, with syncs): {phases}")
        W += 1  # that block is consumed

    e2e = None
    if args.e2e and W + K + K <= n_blocks_needed:
        blk_bytes = sum(p_.numel() * p_.element_size() for p_ in engine.blocks_info["transformers"][W + K].parameters())
        ms_e2e, _ = timed_blocks(W + K, K, host_resident=True)
        log(f"e2e {K} blocks: {ms_e2e / K:.1f} ms/step")
        d2h = sum(b_.numel() * b_.element_size() for b_ in engine.blocks_info["transformers"][W + K].buffers())
        e2e = dict(value=round(TOKENS / (ms_e2e / K * LAYERS / 1000.0), 1), unit="calib tokens/s",
                   h2d_bytes_per_step=blk_bytes, d2h_bytes_per_step=d2h, ms_per_step=round(ms_e2e / K, 2))

    out = dict(metric="Llama-2-7B GPTQ-INT4-g128 calibration throughput (whole model)", value=round(value, 1),
               unit="calib tokens/s", n_gpus=world, steps=K, warmup=W, ms_per_step=round(ms_step, 2),
               higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f16 activations / f32 GPTQ math / u4 codes",
               data="synthetic", gptq_calib_sec_full_model=round(ms_step * LAYERS / 1000.0, 2),
               config=dict(workload="GPTQ INT4 g128 sym block_size=128 percdamp=.01, Llama-2-7B random-init fp16, "
                                    "128 synthetic calib seqs x 2048 tokens; step = one decoder block of the 32",
                           global_batch=N_SAMPLES, seq_len=SEQ, parallelism=f"calib-dp{world}",
                           l2="inputs >> L2 (2.1 GB activations, 0.4 GB weights per step)"),
               gpu_launches=int(launches), clocks=clocks, e2e=e2e, roofline=roofline, phases_ms=phases)
    if rank == 0 and args.decode:
        del model, engine
        torch.cuda.empty_cache()
        out.update(bench_decode(dev, pk))
        log("decode bench done")
    if rank == 0 and world == 1 and args.cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(sample_layers=1)
        log("cpu baseline done")
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def bench_decode(dev, pk):
    """Llama-2-7B INT4 decode: one token through the 224 packed linears (7 per block x 32), CUDA graph, weights of
    all 32 blocks distinct (3.36 GB >> L2)."""
    from neural_compressor_b200 import ops

    packs = []
    g = torch.Generator().manual_seed(3)
    base = {}
    for name, N, Kd in LINEARS:
        Wt = (torch.randn(N, Kd, generator=g) * 0.02).to(dev)
        r = ops.rtn_quant_pack(Wt, 4, 128, True)
        base[name] = (r["qweight"], r["qzeros"], r["scales"], ops.build_stream_layout(r["qweight"], r["qzeros"], r["scales"],
                                                                                         4, 128, Kd, N))
    for _ in range(LAYERS):
        blk = {k: tuple(t.clone() for t in v) for k, v in base.items()}
        # what modules.SiblingGroup builds for q/k/v and gate/up: the members' strip-major layouts, concatenated
        blk["qkv"] = torch.cat([blk[n][3] for n in ("q", "k", "v")])
        blk["gateup"] = torch.cat([blk[n][3] for n in ("gate", "up")])
        packs.append(blk)
    xh = torch.randn(1, HIDDEN, device=dev, dtype=torch.float16)
    yh = {n: torch.empty(1, N, device=dev, dtype=torch.float16) for n, N, _ in LINEARS}
    yh["qkv"] = torch.empty(1, 3 * HIDDEN, device=dev, dtype=torch.float16)
    yh["gateup"] = torch.empty(1, 2 * INTER, device=dev, dtype=torch.float16)

    def token_fused(flags):
        x = xh
        for blk in packs:
            ops.woq_linear_stream(x, blk["qkv"], None, 4, 128, HIDDEN, 3 * HIDDEN, out_dtype=torch.float16, flags=flags,
                                  out=yh["qkv"])
            ops.woq_linear_stream(x, blk["o"][3], None, 4, 128, HIDDEN, HIDDEN, out_dtype=torch.float16, flags=flags,
                                  out=yh["o"])
            ops.woq_linear_stream(x, blk["gateup"], None, 4, 128, HIDDEN, 2 * INTER, out_dtype=torch.float16, flags=flags,
                                  out=yh["gateup"])
            ops.woq_linear_stream(yh["gateup"][:, :INTER], blk["down"][3], None, 4, 128, INTER, HIDDEN,
                                  out_dtype=torch.float16, flags=flags, out=yh["down"])
            x = yh["down"]

    def token(flags, use_stream):
        x = xh
        for blk in packs:
            for name, N, Kd in LINEARS:
                inp = x if Kd == HIDDEN else yh["up"]
                qw, qz, sc, lay = blk[name]
                if use_stream:
                    ops.woq_linear_stream(inp, lay, None, 4, 128, Kd, N, out_dtype=torch.float16, flags=flags, out=yh[name])
                else:
                    ops.woq_linear(inp, qw, qz, sc, None, 4, 128, Kd, N, out_dtype=torch.float16, flags=flags, out=yh[name])
            x = yh["down"]

    res = {}
    by = sum(N * Kd // 2 + 2 * N * Kd // 128 + N * Kd // 256 + 2 * Kd + 2 * N for _, N, Kd in LINEARS) * LAYERS
    for flags, use_stream, tag in ((0, False, "optimum_layout"), (2, False, "optimum_layout_pdl"),
                                   (0, True, "stream_layout"), (2, True, "stream_layout_pdl"),
                                   (2, None, "stream_layout_pdl_siblings_fused")):
        run = (lambda: token_fused(flags)) if use_stream is None else (lambda: token(flags, use_stream))
        run()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            run()
        for _ in range(3):
            graph.replay()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        s.record()
        for _ in range(20):
            graph.replay()
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 20
        res[tag] = dict(ms_per_token=round(ms, 4), tokens_per_s=round(1000.0 / ms, 1), GBs=round(by / ms / 1e6, 1))
    best = max(res.values(), key=lambda r: r["tokens_per_s"])
    return dict(decode=dict(metric="Llama-2-7B INT4 decode tokens/s (224 WOQ linears, batch 1, CUDA graph)", **best,
                            variants=res, bytes_per_token=by, dequant_gemm_tflops=round(
                                2 * sum(N * Kd for _, N, Kd in LINEARS) * LAYERS / (best["ms_per_token"] * 1e9), 2)),
                roofline_decode=dict(bound="hbm", kernel="woq_gemm_stream_kernel / woq_gemm_mma_kernel (best variant)", achieved=best["GBs"],
                                     peak=pk["hbm_gbs"], unit="GB/s", frac=round(best["GBs"] / pk["hbm_gbs"], 4),
                                     traffic=None, peak_source=pk["source"]))


def syrk_traffic_per_launch(flops_per_launch):
    """DRAM bytes per SYRK launch from the newest committed ncu capture (profiles/rNN_traffic.json), averaged over the
    launches of the timed region (small grid = C 4096, large grid = C 11008).  None when no capture is committed."""
    import glob

    files = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r*_traffic.json")))
    if not files or not flops_per_launch:
        return None
    by_grid = sorted(json.load(open(files[-1]))["dram_bytes_per_launch_by_grid"].values())
    if len(by_grid) < 2:
        return None
    cut = (min(flops_per_launch) + max(flops_per_launch)) / 2
    tot = sum(by_grid[-1] if f > cut else by_grid[0] for f in flops_per_launch)
    return round(tot / len(flops_per_launch))


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_baseline(sample_layers=1):
    """The reference's CPU arithmetic (oracle port) on a bounded sample: GPTQ.add_batch for one 2048-token sequence
    at C=4096 and C=11008, GPTQ.fasterquant on one [4096,4096] layer, pack; extrapolated to the whole model with
    the reference's own structure (7 Hessians per block per sample) -- forwards NOT counted."""
    from oracle import woq_oracle as O

    cores = os.cpu_count()
    torch.set_num_threads(cores)
    g = torch.Generator().manual_seed(0)
    t_add = {}
    for C in (HIDDEN, INTER):
        lay = O.GPTQLayerOracle(8, C, bits=4, sym=True)
        x = torch.randn(1, SEQ, C, generator=g)
        lay.add_batch(x)  # warm
        t0 = time.perf_counter()
        lay.add_batch(x)
        t_add[C] = time.perf_counter() - t0
    # column loop: thousands of tiny torch ops -- it does not scale past ~16 threads (more threads only add
    # fork/join overhead), so cap it there; add_batch above used every core
    fq_threads = min(cores, 16)
    torch.set_num_threads(fq_threads)
    lay = O.GPTQLayerOracle(HIDDEN, HIDDEN, bits=4, sym=True)
    for _ in range(3):
        lay.add_batch(torch.randn(1, SEQ, HIDDEN, generator=g))
    Wt = torch.randn(HIDDEN, HIDDEN, generator=g) * 0.02
    t0 = time.perf_counter()
    r = lay.fasterquant(Wt, blocksize=128, percdamp=0.01, groupsize=128)
    t_fq = time.perf_counter() - t0
    t0 = time.perf_counter()
    codes = O.GPTQLayerOracle.export_codes(r["Q"], r["scale"], r["zero"], 128, True)
    O.pack_optimum(codes, r["scale"], None, 4, 128)
    t_pack = time.perf_counter() - t0
    torch.set_num_threads(cores)
    fq_units = sum(N * Kd * Kd for _, N, Kd in LINEARS) / (HIDDEN**3)
    per_block = N_SAMPLES * (6 * t_add[HIDDEN] + t_add[INTER]) + t_fq * fq_units + t_pack * (sum(N * Kd for _, N, Kd in LINEARS) / HIDDEN**2)
    total = per_block * LAYERS
    return dict(value=round(TOKENS / total, 3), unit="calib tokens/s", cores=cores, kind="port",
                fasterquant_threads=fq_threads,
                sample=f"add_batch 1x{SEQ} tokens @C=4096 ({t_add[HIDDEN]:.3f}s) and @C=11008 ({t_add[INTER]:.3f}s), "
                       f"fasterquant+export+pack of one 4096x4096 layer ({t_fq:.2f}s+{t_pack:.2f}s); extrapolated x128 "
                       f"samples x 7 Hessians x 32 blocks, fasterquant scaled by N*C^2, block forwards not counted",
                est_full_model_sec=round(total, 1))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    vals = []
    for _ in range(max(1, min(args.steps, 2))):  # each step is the bounded sample; a few repeats at most
        vals.append(cpu_baseline())
    b = max(vals, key=lambda v: v["value"])
    out = dict(impl="reference", metric="Llama-2-7B GPTQ-INT4-g128 calibration throughput (whole model)",
               value=b["value"], unit="calib tokens/s", n_gpus=int(os.environ.get("WORLD_SIZE", "1")), steps=args.steps,
               warmup=args.warmup, ms_per_step=round(b["est_full_model_sec"] * 1000 / LAYERS, 1), higher_is_better=True,
               scaling="strong", vs_baseline=None, dtype="f32", data="synthetic",
               config=dict(workload="GPTQ INT4 g128 sym block_size=128, Llama-2-7B shapes, 128x2048 calib tokens; "
                                    "CPU arithmetic of the reference on a bounded sample, extrapolated"),
               cpu_baseline=b, e2e=dict(value=b["value"], unit="calib tokens/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
               gpu_launches=0, wall_s=round(time.perf_counter() - t0, 1))
    print(json.dumps(out))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-e2e", dest="e2e", action="store_false")
    ap.add_argument("--phases", action="store_true", help="also report a per-phase breakdown of one extra block")
    ap.add_argument("--no-decode", dest="decode", action="store_false")
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
