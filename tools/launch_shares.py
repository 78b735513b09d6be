"""profiles/r02_launch_shares.md from the ncu launch list of bench.py and the bench line itself."""
import collections
import csv
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = list(csv.reader(open(os.path.join(ROOT, "profiles", "r02_bench_launches.csv"))))
hdr, agg = None, collections.OrderedDict()
for r in rows:
    if hdr is None:
        if r and r[0] == "ID":
            hdr = r
        continue
    if len(r) < len(hdr):
        continue
    d = dict(zip(hdr, r))
    name = d["Kernel Name"].split("(")[0].replace("void ", "")
    v = float(d["Metric Value"].replace(",", ""))
    v = v / 1e3 if d["Metric Unit"] == "ns" else (v * 1e3 if d["Metric Unit"] == "ms" else v)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
step = ["inner_maximize_kernel<4, 2, 1, 10, 1>", "grad_vector_b_staged_kernel<4, 2>", "trsm_kernel<1>",
        "trsm_kernel<0>", "cross_cov_kernel", "candidate_setup_kernel", "acq_reduce_kernel"]
tot = sum(agg[k][1] for k in step if k in agg)
b = json.loads(open(os.path.join(ROOT, "profiles", "r02_bench_1gpu.json")).readline())
out = ["# Round 2 — kernel shares of a C3 step: ncu launch list against the bench line", "",
       "`r02_bench_launches.csv`: `ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv python "
       "bench.py --steps 2",
       "--warmup 1 --no-c5 --no-cpu-baseline` (per-launch times under ncu are cold-cache and serialised; the "
       "SHARES are what is",
       "compared).  Kernels of the timed step (stage B, stage C), summed over the 9 steps the command runs "
       "(warm-up, timed, the two",
       "end-to-end legs):", "", "| kernel | launches | total under ncu | share of the step kernels |",
       "|---|---|---|---|"]
for k in step:
    if k in agg:
        out.append("| `%s` | %d | %.1f ms | %.1f %% |" % (k, agg[k][0], agg[k][1] / 1e3, 100 * agg[k][1] / tot))
out += ["", "Bench line (`r02_bench_1gpu.json`, CUDA events inside `bench.py`, no profiler): `roofline.kernel_ms` = "
        "%.1f of `ms_per_step` = %.1f," % (b["roofline"]["kernel_ms"], b["ms_per_step"]),
        "`kernel_share_of_step` = %.3f.  The shares agree (the step also holds launch gaps, the L2-flush fill and "
        "the argmax," % b["roofline"]["kernel_share_of_step"],
        "which the launch list does not see).", "",
        "Everything else in the list belongs to the set-up outside the timed steps: the FP64 probes "
        "(`dfma_probe_kernel`,",
        "`dmma_probe_kernel`), the cuBLAS DGEMM anchor (`cutlass_80_tensorop_d884gemm`), stage A of the 20 "
        "hyper-samples",
        "(`potrf_persistent_kernel`, `inv_syrk_contract_kernel`, `assemble_kernel`, ...) and the W-distance tables.",
        "", "Regenerate with `python tools/launch_shares.py`."]
open(os.path.join(ROOT, "profiles", "r02_launch_shares.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out[8:20]))
