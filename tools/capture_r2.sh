#!/bin/bash
# Runs on the GPU box (gpurun): ncu captures behind profiles/*_r2_*.  Numbers printed by runs under ncu are never bench values.
set -u
OUT=gpurun_out
B="python bench.py --steps 1 --warmup 1 --stages 0 --cpu-baseline 0 --multi 0 --torch-cuda-baseline 0 --second-head 0 --stream-probe 0"
# (a) launch list of one inference step (+ the warm-up step and the overlap-off pass): per-launch durations
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches_r2.csv $B > $OUT/launches_r2.log 2>&1
# (b) full captures of one chunk's kernels of the anchor phase (4th chunk of the 2nd inference of tools/ncu_targets.py)
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"CoarseEpi|xw_gemm|xw_head|xw_cand|xw_cell|gather_anchor" \
  --launch-skip 18 --launch-count 6 -f -o $OUT/ncu_r2_anchor python tools/ncu_targets.py infer > $OUT/ncu_r2_anchor.log 2>&1
# (c) every kernel of one ViT-L block on one frame: time + tensor-pipe activity
timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none \
  --csv --log-file $OUT/vit_block_r2.csv python tools/ncu_targets.py vit > $OUT/vit_block_r2.log 2>&1
# (d) full captures of the attention kernel and the fc1 + GELU GEMM of that block
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"flash_attn|EpiGelu" --launch-skip 2 --launch-count 2 -f \
  -o $OUT/ncu_r2_vit python tools/ncu_targets.py vit > $OUT/ncu_r2_vit.log 2>&1
ls -la $OUT/*.ncu-rep $OUT/*.csv
# ---- final state of the round (profiles/ncu_r2_final_*.csv, launches_r2_summary.csv) ----
# (e) exact box GEMM (tokens on UMMA M), window extraction and head of one chunk
timeout 500 ncu --set full --import-source on --clock-control none -k regex:"xw_gemm|xw_head|xw_window" --launch-skip 24 --launch-count 3 \
  -f -o $OUT/ncu_r2_final_xw python tools/ncu_targets.py infer > $OUT/nf1.log 2>&1
# (f) the GEMMs and the attention kernel of one ViT-L block (P through tensor memory, 25 % polynomial exp2)
timeout 400 ncu --set full --import-source on --clock-control none -k regex:"flash_attn|EpiResidual|tc_gemm2" --launch-skip 5 --launch-count 5 \
  -f -o $OUT/ncu_r2_final_vit python tools/ncu_targets.py vit > $OUT/nf2.log 2>&1
# (g) launch list of one step of the final code -> python tools/launch_summary.py $OUT/launches_r2_final.csv profiles/launches_r2_summary.csv
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches_r2_final.csv $B > $OUT/launches_r2_final.log 2>&1
# summaries: python tools/ncu_summary.py $OUT/<report>.ncu-rep profiles/<name>.csv
