#!/bin/bash
# Round evidence (run under gpurun): launch list of a bench step + one `ncu --set full` capture per kernel class.
# Every ncu command runs only after the same command has exited 0 without ncu.  Outputs land in gpurun_out/.
mkdir -p gpurun_out
python -c "import torch" > /dev/null 2>&1
B="python bench.py --steps 1 --warmup 1 --windows 16 --no-cpu-baseline --no-e2e"
timeout -s KILL 300 $B > gpurun_out/plain_bench_w16.log 2>&1 && \
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
for t in mel mel-fused attn ln gemm-fc1 gemm-qkv gemm-o gemm-fc2; do
  case $t in mel|mel-fused) k="logmel_tiles";; attn) k="attention_fwd";; ln) k="layernorm";; *) k="gemm_bf16";; esac
  timeout -s KILL 120 python tests/prof_target.py $t > gpurun_out/plain_$t.log 2>&1 && \
  timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:"$k" -s 2 -c 1 -o gpurun_out/final_$t -f python tests/prof_target.py $t > gpurun_out/ncu_final_$t.log 2>&1
  echo "ncu $t exit $?"
done
# row f1: launch lists of a few decode steps (1 and 64 windows) and one --set full capture per decode kernel class
for B in 64 1; do
  timeout -s KILL 120 python tests/gpu_diag_decode.py prof $B > gpurun_out/prof_plain_$B.log 2>&1 && \
  timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/dec_launches_b$B.csv python tests/gpu_diag_decode.py prof $B > gpurun_out/ncu_dec_$B.log 2>&1
  echo "decode launch list B=$B exit $?"
done
for t in dec-xattn dec-skinny dec-logits; do
  case $t in dec-xattn) k="decode_attention";; *) k="skinny_gemm";; esac
  timeout -s KILL 120 python tests/prof_target.py $t > gpurun_out/plain_$t.log 2>&1 && \
  timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:"$k" -s 2 -c 1 -o gpurun_out/final_$t -f python tests/prof_target.py $t > gpurun_out/ncu_final_$t.log 2>&1
  echo "ncu $t exit $?"
done
timeout -s KILL 400 python tests/gpu_diag_decode.py perf 1 8 64 > gpurun_out/dec_perf.log 2>&1; grep "graph+pdl" gpurun_out/dec_perf.log
timeout -s KILL 400 python bench.py --mode decode --steps 3 --warmup 1 > gpurun_out/bench_decode_final.log 2> gpurun_out/bench_decode_final.err; echo "bench decode final exit $?"
timeout -s KILL 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_final.log 2> gpurun_out/bench_final.err; echo "bench final exit $?"
timeout -s KILL 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_final.log 2> gpurun_out/bench_ref_final.err; echo "bench ref exit $?"
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -n 2 gpurun_out/smoke.log
