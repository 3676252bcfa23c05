"""Turns the ncu outputs in gpurun_out/ (tests/gpu_profile.sh) into the committed evidence under profiles/<round>/:
launch-list summary, one text summary per `--set full` capture (key raw metrics + top stall sites), and
profiles/traffic.json (dram bytes per launch, read by bench.py for `roofline.traffic`).

    python tests/make_profiles.py r01"""
import csv
import json
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rnd = sys.argv[1] if len(sys.argv) > 1 else "r01"
out_dir = os.path.join(ROOT, "profiles", rnd)
os.makedirs(out_dir, exist_ok=True)
G = os.path.join(ROOT, "gpurun_out")

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg"]
UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}


def raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    return {h: (u, v) for h, u, v in zip(rows[0], rows[1], rows[2])}


def source_summary(rep, top=14):
    path = "/tmp/_src.csv"
    with open(path, "w") as f:
        f.write(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout)
    return subprocess.run([sys.executable, os.path.join(ROOT, "tests", "ncu_src_summary.py"), path, str(top)],
                          capture_output=True, text=True).stdout


traffic = {}
names = {"mel": "logmel_tiles_kernel<false> (FeatureExtractor.__call__: f32 [n_mels, frames] output)",
         "mel-fused": "logmel_tiles_kernel<true> (aries_encode_pcm: PCM -> bf16 time-major conv1 operand, 64 windows, 128 bins)",
         "attn": "attention_fwd_kernel", "ln": "layernorm_kernel (ln_post: the one LayerNorm launch left per forward pass)",
         "gemm-fc1": "gemm_bf16_tcgen05 (fc1: M=96000 N=5120 K=1280, f16 x f16, LayerNorm-folded + GELU epilogue)",
         "gemm-qkv": "gemm_bf16_tcgen05 (QKV: M=96000 N=3840 K=1280, f16 x f16, LayerNorm-folded epilogue, Q|K row-major + V transposed)",
         "gemm-o": "gemm_bf16_tcgen05 (out-proj: M=96000 N=1280 K=1280, bias + f16 residual epilogue + LayerNorm partials)",
         "gemm-fc2": "gemm_bf16_tcgen05 (fc2: M=96000 N=1280 K=5120, bias + f16 residual epilogue + LayerNorm partials)",
         "dec-xattn": "decode_attention_kernel (row f1: cross-attention of one decode step, 64 windows x 20 heads x 1500 keys)",
         "dec-skinny": "skinny_gemm_tcgen05 (row f1: fc1 of one decode step, 64 sequences, N=5120 K=1280, cluster split-K 7)",
         "dec-logits": "skinny_gemm_tcgen05 (row f1: logits of one decode step, 64 sequences, N=51866 K=1280)"}
for tag, name in names.items():
    rep = os.path.join(G, f"final_{tag}.ncu-rep")
    if not os.path.exists(rep):
        continue
    m = raw(rep)
    lines = [f"# ncu --set full --clock-control none --import-source on: {name}",
             f"# target: python tests/prof_target.py {tag}   (launch 3 of the target; see tests/gpu_profile.sh)", ""]
    for k in KEYS:
        hit = [h for h in m if h == k or h.startswith(k + " ")]
        for h in hit[:1]:
            lines.append(f"{h:75s} {m[h][1]:>16s} {m[h][0]}")
    stalls = sorted(((float(v[1]), h) for h, v in m.items()
                     if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")
                     and v[1] not in ("", "n/a")), reverse=True)[:6]
    lines.append("")
    lines.append("top stall reasons (warps stalled per issue-active cycle): " +
                 ", ".join(f"{h.split('stalled_')[1].split('_per_')[0]}={x:.2f}" for x, h in stalls))
    rd = float(m["dram__bytes_read.sum"][1]) * UNIT[m["dram__bytes_read.sum"][0]]
    wr = float(m["dram__bytes_write.sum"][1]) * UNIT[m["dram__bytes_write.sum"][0]]
    lines.append(f"dram bytes per launch: read {rd / 1e6:.1f} MB + write {wr / 1e6:.1f} MB = {(rd + wr) / 1e6:.1f} MB")
    traffic[tag] = {"kernel": name, "dram_read_bytes": rd, "dram_write_bytes": wr, "dram_bytes": rd + wr,
                    "duration_us_under_ncu": float(m["gpu__time_duration.sum"][1])}

    def pct(key):
        hit = [h for h in m if h == key or h.startswith(key + " ")]
        try:
            return float(m[hit[0]][1]) / 100.0 if hit else None
        except ValueError:
            return None
    traffic[tag]["fma_pipe_frac"] = pct("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed")
    traffic[tag]["issue_active_frac"] = pct("sm__issue_active.avg.pct_of_peak_sustained_elapsed")
    traffic[tag]["tensor_pipe_frac"] = pct("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed")
    traffic[tag]["xu_pipe_frac"] = pct("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed")
    lines += ["", "## source page (sampled stalls)", source_summary(rep)]
    with open(os.path.join(out_dir, f"ncu_full_{tag}.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")

# launch list -> per-kernel totals
lc = os.path.join(G, "launches.csv")
if os.path.exists(lc):
    rows = list(csv.reader(l for l in open(lc) if l.startswith('"')))
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows[1:]:
        if len(r) != len(hdr):
            continue
        tot[r[ki]] += float(r[vi].replace(",", ""))
        cnt[r[ki]] += 1
    unit = rows[1][hdr.index("Metric Unit")]
    scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3}.get(unit, 1.0)
    total = sum(tot.values())
    with open(os.path.join(out_dir, "launches_bench_w16_summary.md"), "w") as f:
        f.write("# ncu launch list — `python bench.py --steps 1 --warmup 1 --windows 16 --no-cpu-baseline --no-e2e`\n\n"
                "`ncu --metrics gpu__time_duration.sum --clock-control none -c 700` (cold-cache, serialised: compare "
                "SHARES with bench.py's CUDA-event `kernels` object, not absolute times). Covers the warm-up step and "
                "the timed step plus torch's own fill/copy kernels.\n\n| kernel | launches | total us | share |\n|---|---|---|---|\n")
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
            f.write(f"| `{k[:70]}` | {cnt[k]} | {v * scale:.1f} | {100 * v / total:.1f}% |\n")
    import shutil
    shutil.copy(lc, os.path.join(out_dir, "launches_bench_w16.csv"))

# row f1: launch lists of a few decode steps (tests/gpu_diag_decode.py prof B)
for B in (1, 64):
    lc = os.path.join(G, f"dec_launches_b{B}.csv")
    if os.path.exists(lc):
        title = (f"ncu launch list -- `python tests/gpu_diag_decode.py prof {B}`: 6 greedy decode steps, {B} window(s), large-v3 "
                 "layer shape with 4 layers, direct stream launches (no graph, no programmatic launch)")
        md = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "summarize_launches.py"), lc, title],
                            capture_output=True, text=True).stdout
        with open(os.path.join(out_dir, f"launches_decode_b{B}_summary.md"), "w") as f:
            f.write(md)

# bench.py reads the fc1 GEMM as the representative launch of the dominant kernel class
tj = {"source": f"profiles/{rnd}/ncu_full_*.txt (ncu --set full, one launch each, tests/prof_target.py shapes)",
      "gemm_bf16_tcgen05": traffic.get("gemm-fc1", {}).get("dram_bytes"),
      "logmel_tiles_kernel": traffic.get("mel-fused", traffic.get("mel", {})).get("dram_bytes"),
      "decode_attention_kernel": traffic.get("dec-xattn", {}).get("dram_bytes"),
      "logmel_tiles_kernel_ncu": {"fma_pipe_frac": traffic.get("mel-fused", {}).get("fma_pipe_frac"),
                                  "issue_active_frac": traffic.get("mel-fused", {}).get("issue_active_frac"),
                                  "source": f"profiles/{rnd}/ncu_full_mel-fused.txt"},
      "detail": traffic}
json.dump(tj, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps({k: round(v["dram_bytes"] / 1e6, 1) for k, v in traffic.items()}))
