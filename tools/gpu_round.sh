#!/bin/bash
# One GPU-box visit: parity tests, the bench line (both arms), the ncu launch list and one --set full capture.
# Usage (from the repo root, under gpurun):  bash tools/gpu_round.sh [tag]
# Everything lands in gpurun_out/<tag>_*; numbers printed by the runs under ncu are never bench values.
tag=${1:-r1}
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
export CB_REQUIRE_GPU=1
K='regex:conv1|dwpw|dw_kernel|pw_gemm|vlad|scores|topk|finalize|merge_lists|split_|dls_|pnp_|icp_|copy_models|hamming|gms_|collect_'

timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a $out/${tag}_pytest_gpu.log
tail -3 $out/${tag}_pytest_gpu.log

timeout 300 python __graft_entry__.py --smoke > $out/${tag}_smoke.log 2>&1
echo "smoke exit $?"; tail -2 $out/${tag}_smoke.log

timeout 420 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err
echo "bench exit $?"; cat $out/${tag}_bench.json

timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err
echo "reference arm exit $?"; cat $out/${tag}_bench_reference.json

timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/${tag}_launches_run.log 2>&1
echo "ncu launch list exit $?"

timeout 600 ncu --set full --clock-control none --import-source on -k "$K" -s 40 -c 34 -f -o $out/${tag}_prof \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/${tag}_prof_run.log 2>&1
echo "ncu full exit $?"
if [ -f $out/${tag}_prof.ncu-rep ]; then
  ncu -i $out/${tag}_prof.ncu-rep --page raw --csv > $out/${tag}_prof_raw.csv 2>/dev/null
  ls -la $out/${tag}_prof.ncu-rep
  # the report itself may exceed the 64 MiB transfer limit; the raw page is what gets committed
  sz=$(stat -c %s $out/${tag}_prof.ncu-rep); if [ "$sz" -gt 40000000 ]; then rm $out/${tag}_prof.ncu-rep; fi
fi
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.draw --format=csv > $out/${tag}_gpu.txt

# the other two single-GPU configurations of BASELINE.json and the component micro-benchmarks
timeout 300 python bench.py --workload config2 --no-cpu-baseline > $out/${tag}_bench_config2.json 2> $out/${tag}_bench_config2.err
echo "config2 exit $?"
timeout 400 python bench.py --workload pnp5 --steps 3 --warmup 1 > $out/${tag}_bench_pnp5.json 2> $out/${tag}_bench_pnp5.err
echo "pnp5 exit $?"
timeout 200 python tools/bench_pnp.py > $out/${tag}_pnp.jsonl 2>&1
timeout 200 python tools/bench_features.py > $out/${tag}_features.json 2>/dev/null
timeout 200 python tools/bench_frontend.py > $out/${tag}_frontend.json 2>/dev/null
timeout 200 python tools/bench_desc.py 64 --layers > $out/${tag}_desc_layers.json 2>/dev/null
# front-end kernels under ncu --set full (remap, ORB, StereoBM, matcher, GMS)
FK='regex:sbm_|fast_kernel|describe_kernel|harris_kernel|blur_|resize_kernel|compact_kernel|scan_rows|remap_kernel|hamming|gms_|expand_bits|unpack_matches'
timeout 400 ncu --set full --clock-control none -k "$FK" -c 45 -f -o $out/${tag}_fe_prof python tools/bench_features.py > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none -k "$FK" -c 12 -f -o $out/${tag}_fe2_prof python tools/bench_frontend.py > /dev/null 2>&1
for f in fe fe2; do
  if [ -f $out/${tag}_${f}_prof.ncu-rep ]; then
    ncu -i $out/${tag}_${f}_prof.ncu-rep --page raw --csv > $out/${tag}_${f}_prof_raw.csv 2>/dev/null
    rm $out/${tag}_${f}_prof.ncu-rep
  fi
done
echo "front-end ncu done"
