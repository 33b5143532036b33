#!/bin/bash
O=gpurun_out/r03t; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_model.py -x -q -m gpu > $O/pytest.log 2>&1; tail -3 $O/pytest.log
for v in pre nopre pre nopre; do
if [ $v = nopre ]; then export FRTM_B200_LIB=$PWD/frtm_vos_b200/libfrtm_b200_timing.so; else unset FRTM_B200_LIB; fi
timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > $O/bench2_$v.json 2> $O/bench2.err
echo "$v"; python tools/bench_brief.py $O/bench2_$v.json 2>&1 | head -1 | cut -c1-90
done
for v in pre nopre; do
if [ $v = nopre ]; then export FRTM_B200_LIB=$PWD/frtm_vos_b200/libfrtm_b200_timing.so; else unset FRTM_B200_LIB; fi
timeout 900 python bench.py --config 3 --steps 4 --warmup 3 --no-cpu-baseline > $O/bench3_$v.json 2> $O/bench3.err
echo "$v"; python tools/bench_brief.py $O/bench3_$v.json 2>&1 | head -1 | cut -c1-90
done
