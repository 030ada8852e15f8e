"""Assemble the round-2 profile documents under profiles/ from the final capture (gpurun_out/r02f, tools/r02_final_capture.sh
plus the 8-GPU runs).  Usage: python tools/make_r02_profiles.py"""
import json
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "gpurun_out", "r02f")
OUT = os.path.join(ROOT, "profiles")


def load(name):
    with open(os.path.join(SRC, name)) as f:
        return json.loads(f.read().strip().splitlines()[-1])


def text(name):
    with open(os.path.join(SRC, name)) as f:
        return f.read()


def bench_lines():
    rows = [("bench.py (default: config 3 shape, B=8, N=512, paper dims)", "bench_default.json"),
            ("bench.py --workload config1", "bench_config1.json"), ("bench.py --workload config2", "bench_config2.json"),
            ("bench.py --workload config5", "bench_config5.json"),
            ("bench.py --workload train (TrainStepGraph, default)", "bench_train_graph.json"),
            ("bench.py --workload train --train-mode manual", "bench_train_manual.json"),
            ("bench.py --workload train --train-mode autograd", "bench_train_autograd.json"),
            ("bench.py --workload train --train-sizes 2 (8+366, 20+250 tokens: N = 374)", "bench_train_n374.json"),
            ("torchrun x8: bench.py --gpus 8", "bench_default_8gpu.json"),
            ("torchrun x8: bench.py --gpus 8 --workload train", "bench_train_8gpu.json")]
    out = ["# bench.py lines of round 2, final build (one fresh B200 per gpurun call; tools/r02_final_capture.sh, tools/make_r02_profiles.py)",
           "", "| command | metric | value | ms / step | launches / step |", "|---|---|---:|---:|---:|"]
    full = []
    for cmd, f in rows:
        d = load(f)
        out.append(f"| `{cmd}` | {d['metric']} | {d['value']:.3f} {d['unit']} | {d['ms_per_step']:.2f} | {d.get('gpu_launches_per_step', '-')} |")
        full.append(json.dumps(d))
    d = load("bench_default.json")
    out += ["", "`--impl reference` (CPU oracle port, full batch of 8, 16 host cores; measured with the mid-round build, the CPU path did not change): "
            "0.021 steps/s, 47.6 s per step.", "", "Blocks of the default line:", "", "```json"]
    for k in ("e2e", "sustained", "ragged", "sample_parallel", "roofline", "gpu_eager_baseline", "cpu_baseline", "clocks"):
        out.append(json.dumps({k: d[k]}))
    out += ["```", "", "Full JSON lines:", "", "```json"] + full + ["```", ""]
    open(os.path.join(OUT, "r02_bench_lines.md"), "w").write("\n".join(out))


def launches():
    d = load("bench_default.json")
    body = text("launches_step.md").splitlines()
    head = [f"# ncu launch list of one sampling step, final build of round 2: 115 launches, {body[0].split('launches, ')[1]}",
            "",
            "Command (one B200, gpurun, tools/r02_final_capture.sh): `ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file",
            "gpurun_out/r02f/launches_step.csv python bench.py --steps 1 --warmup 3 --profile-eager`, then `python tools/summarize_launches.py ... 115`",
            "(raw list: profiles/r02_launches.csv).  Per-launch times are cold-cache and serialised; the same step replayed as a CUDA graph takes",
            f"{d['ms_per_step']:.2f} ms on the same box (bench.py), so the SHARES are what carries over.  Attention core: 38.7 % here, 8 x "
            f"{d['roofline']['ms_per_launch']:.3f} ms / {d['ms_per_step']:.2f} ms = {800 * d['roofline']['ms_per_launch'] / d['ms_per_step']:.1f} % in the timed step.",
            "",
            f"Round 1 -> round 2 (same shape): 119 -> 115 launches, 28.65 -> {body[0].split('launches, ')[1].split(' ms')[0]} ms serialised, 27.52 -> {d['ms_per_step']:.2f} ms per graph replay.",
            "`triattn_flash_g4_kernel` 1.365 -> 1.27 (first key tile on the fast path, overflow check through the row sum), `pair_transition_ws`",
            "0.479 -> 0.388 alone (rows by TMA, hi half of W2) / 0.48 as the average of the 3 launches that also emit the next block's attention bias",
            "and the one that does not, `pair_bias_kernel` 5 x 0.125 -> 1 x 0.158, `gemm_f16_kernel<256, 4>` 0.089 -> 0.083 and `<128, 5>` 0.032 -> 0.023",
            "(eight epilogue warps, accumulator chunk in registers instead of a local-memory array), `trimul_in_t` 0.360 -> 0.326 and `triattn_proj`",
            "0.339 -> 0.299 (one thread per row normalises), `triattn_out` 0.240 -> 0.216, `pair_embed_lut` 0.536 -> 0.444, `coord_head` 0.307 -> 0.225",
            "(row tiles by TMA instead of per-thread bulk copies; `trimul_out` shows 0.36 here both ways -- cold cache, serialised -- and 0.45 ms of the",
            "replayed step in the `PRD_ROW_TMA` A/B): DESIGN section 4.", ""]
    open(os.path.join(OUT, "r02_launches.md"), "w").write("\n".join(head + body[2:]) + "\n")
    shutil.copy(os.path.join(SRC, "launches_step.csv"), os.path.join(OUT, "r02_launches.csv"))


def train():
    g = load("bench_train_graph.json")
    out = ["# Training step (BASELINE config 4) of the final build of round 2: per-op and per-kernel GPU time",
           "",
           f"B = 2, N = 314 (24 + 290 and 35 + 212 tokens), paper dims, one B200; `bench.py --workload train`: {g['ms_per_step']:.2f} ms per step "
           f"({g['value']:.1f} steps/s, {g['gpu_launches_per_step']} launches of libprd_sm100.so per step).",
           "",
           "## How the step got from 71.9 ms to 31.4 ms (each line: one commit, same workload, A/B on one box each)",
           "",
           "| change | ms / step |", "|---|---:|",
           "| mid-round build (fp32 SIMT attention, scalar elementwise kernels, four epilogue warps) | 71.6 |",
           "| four-group forward attention core for every N <= 2048 (was N % 512 == 0 only) | 70.3 |",
           "| backward attention on the tensor cores (mma.sync tf32: forward recompute, dq, dk / dv) | 55.3 |",
           "| GEMM epilogue: eight epilogue warps with their own register budget, TMA tile stores, accumulator chunk in registers (the dynamically indexed `v[32]` lived in local memory), two-instruction tf32 rounding | 48.2 |",
           "| attention kernels: log2-domain scores, one-instruction operand rounding, key state as a min() operand | 46.2 |",
           "| elementwise / LayerNorm kernels vectorised (float4 rows, 32-bit index math instead of three 64-bit divisions per element), transposed channel planes straight from the rows | 41.6 |",
           "| rows <-> channel planes: 64 x 64 tiles, float4 rows | 40.2 |",
           "| weight preparation of every op in one launch (1194 -> 1019 launches), next batch prepared on a side stream | 39.6 |",
           "| bias gradients added up inside the dW kernel from the tiles it stages | 38.8 |",
           "| dW partial sums leave as whole 128-byte rows (transposed through shared memory) instead of 32 scattered atomics per instruction | 36.6 |",
           "| the forward keeps the input of every residual update (1.2 GB) instead of one checkpoint per block + a re-run of the block (`PRD_TAPE=blocks` A/B: 37.2 vs 35.1 on one box; measured after the captures below) | 35.1 |",
           "| the split-operand recompute of every ReLU layer as ONE tf32 GEMM over K = 3 C_in with bias + ReLU + rounding in its epilogue (was three GEMMs chained through memory + a ReLU pass; two builds A/B on one box: 32.6 vs 31.4, 759 vs 729 launches) | 31.4 |",
           "",
           "## Per op (tools/train_breakdown.py: CUDA events around every C-ABI call, serialised)", "", "```", text("train_breakdown.txt").strip(), "```", "",
           "## Per kernel (tools/train_breakdown.py --kernels: torch.profiler / CUPTI over one step)", "", "```",
           "\n".join(l for l in text("train_kernels.txt").splitlines() if "Warn" not in l and "_warn" not in l).strip(), "```", "",
           "## ncu --set full of the backward kernels (gpurun_out/r02f/bwd.ncu-rep, `--clock-control none`, B = 2, N = 314)", "",
           "| kernel | us | dram read / write MB | issue active | tensor pipe | XU pipe |", "|---|---:|---|---:|---:|---:|"]
    cur, rows = None, []
    for line in text("bwd_digest.txt").splitlines():
        if line.startswith("=="):
            cur = {"name": line[3:].split("(")[0].replace("void unnamed>::", "").strip()}
            rows.append(cur)
        elif cur is not None:
            p = line.split()
            if len(p) >= 2:
                cur[p[0]] = p[1]
    seen = {}
    for r in rows:
        key = (r["name"], round(float(r.get("dram__bytes_read.sum", 0)) / 50))
        if key in seen:
            continue
        seen[key] = 1
        out.append(f"| `{r['name']}` | {float(r.get('gpu__time_duration.sum', 0)):.0f} | {float(r.get('dram__bytes_read.sum', 0)):.0f} / "
                   f"{float(r.get('dram__bytes_write.sum', 0)):.0f} | {float(r.get('smsp__issue_active.avg.pct_of_peak_sustained_active', 0)):.0f} % | "
                   f"{float(r.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0)):.0f} % | "
                   f"{float(r.get('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 0)):.0f} % |")
    out += ["",
            "The attention kernels (mma.sync m16n8k8 tf32, `HMMA.1688.F32.TF32`) run the legacy tensor pipe 23 - 38 % and the issue slots 40 - 46 % busy with",
            "16 warps per SM: ~80 instructions per 32 x 8 score tile, of which 12 are MMAs.  The dW kernel reads exactly its operands (101 MB for",
            "a 64 x 64 gradient over 197 k rows) at 4.0 - 5.5 TB/s.", ""]
    open(os.path.join(OUT, "r02_train_breakdown.md"), "w").write("\n".join(out))


def gemm():
    out = ["# tf32 GEMM / dW shapes of the backward pass, final build of round 2 (tools/gemm_shapes.py, one B200, timed alone, L2 warm)",
           "",
           "Before = the mid-round build (four epilogue warps, `float v[32]` indexed dynamically -> local memory, cvt.rna emulation, per-lane",
           "row atomics in the dW epilogue).  R = 197192 rows = B 2 x N 314 x N 314.",
           "",
           "| shape | before ms | after ms |", "|---|---:|---:|",
           "| pre = p Wcat^T, [R, 64] -> [R, 320] | 0.162 | 0.074 |", "| qkvg = x Wcat^T, [R, 64] -> [R, 256] | 0.112 | 0.054 |",
           "| o = xn Wo^T, [R, 64] -> [R, 64] | 0.034 | 0.023 |", "| dh = dy W2^T with the ReLU gate operand, [R, 64] -> [R, 256] | 0.467 | 0.124 |",
           "| plane GEMM, 128 planes of 316 x 316 x 316 | 0.070 | 0.050 |", "| dW 64 x 64 over R rows (+ bias gradient) | 0.064 | 0.019 |",
           "| dW 256 x 64 | 0.067 | 0.045 |", "| dW 64 x 256 | 0.095 | 0.051 |", "", "```", text("gemm_shapes.txt").strip(), "```", ""]
    open(os.path.join(OUT, "r02_gemm_shapes.md"), "w").write("\n".join(out))


if __name__ == "__main__":
    bench_lines()
    launches()
    train()
    gemm()
    print("profiles written")
