#!/usr/bin/env python3
"""profiles/step_counters.json from an ncu CSV over EVERY launch of one step of the bench workload:

    ncu --metrics smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,\\
smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,\\
dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none -k regex:'k_' -c 1300 --csv \\
        --log-file gpurun_out/rNNx_step_counters.csv python tools/prof_render.py standard bdpt 16
    python tools/ncu_step_counters.py gpurun_out/rNNx_step_counters.csv profiles/step_counters.json

Per kernel and for the whole step: launches, warp instructions, thread instructions, fp32 flops (fadd + fmul + 2 ffma),
DRAM bytes, L2 bytes — sums over the step.  bench.py divides them by the live step / kernel times."""
import collections
import csv
import json
import sys

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 10]
hdr = rows[0]
iI, iK, iM, iV, iU = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
field = {"smsp__inst_executed.sum": ("warp_inst", 1), "smsp__thread_inst_executed.sum": ("thread_inst", 1),
         "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum": ("fp32_flops", 1),
         "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum": ("fp32_flops", 1),
         "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum": ("fp32_flops", 2),
         "dram__bytes_read.sum": ("dram_bytes", 1), "dram__bytes_write.sum": ("dram_bytes", 1), "lts__t_bytes.sum": ("l2_bytes", 1)}
agg = collections.defaultdict(lambda: collections.defaultdict(float))
ids = collections.defaultdict(set)
for r in rows[1:]:
    if r[iM] not in field:
        continue
    name = r[iK].split("(")[0].replace("<unnamed>::", "").replace("void ", "").split("<")[0]    # k_path<1> (a template instance) -> k_path
    key, w = field[r[iM]]
    agg[name][key] += w * float(r[iV].replace(",", "")) * scale.get(r[iU], 1)
    ids[name].add(r[iI])
kernels = {}
step = collections.defaultdict(float)
for name, a in agg.items():
    kernels[name] = {"launches": len(ids[name]), **{k: a[k] for k in ("warp_inst", "thread_inst", "fp32_flops", "dram_bytes", "l2_bytes")}}
    for k, v in a.items():
        step[k] += v
    step["launches"] += len(ids[name])
out = {"source": "ncu over every launch of one Cornell-Standard 784x784 BDPT 16 spp frame (" + sys.argv[1].split("/")[-1] + ")",
       "step": dict(step), "kernels": kernels}
json.dump(out, open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout, indent=1)
print("step: %.1f G warp instructions, %.2f lanes/instruction, %.1f GFLOP fp32, %.1f GB DRAM, %.1f GB L2 over %d launches" %
      (step["warp_inst"] / 1e9, step["thread_inst"] / max(1.0, step["warp_inst"]), step["fp32_flops"] / 1e9, step["dram_bytes"] / 1e9,
       step["l2_bytes"] / 1e9, int(step["launches"])), file=sys.stderr)
