#!/bin/bash
# Round-end evidence on the GPU box: bench lines, smoke, launch list, DRAM traffic of every launch of the default workload,
# ncu --set full of one bounce of each kernel (+ the F32 extend kernel).
# usage: TAG=r1c bash tools/final_capture.sh     (outputs under gpurun_out/)
TAG=${TAG:-final}
O=gpurun_out
python -c "import bench; print(bench.source_sha256())" > $O/lib_sha_$TAG.txt   # ties the ncu figures to these sources (tools/kernel_traffic.py)
sha256sum craytracer_b200/libcray_b200.so >> $O/lib_sha_$TAG.txt
python bench.py > $O/bench_$TAG.json 2> $O/bench_$TAG.err
python bench.py --impl reference > $O/bench_${TAG}_reference.json 2> $O/bench_${TAG}_reference.err
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$TAG.log 2>&1
SHORT="python bench.py --spp 256 --steps 1 --warmup 1 --no-cpu-baseline --no-f32-leg"
$SHORT > $O/plain_$TAG.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file $O/launches_$TAG.csv $SHORT > $O/ncu_a_$TAG.log 2>&1
# every launch of one default-size render, twice (device-film and host-film leg, same seed): tools/kernel_traffic.py halves it
FULL="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-f32-leg"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'k_begin|k_generate|k_requeue|k_wide_persistent|k_shade' --csv --log-file $O/traffic_$TAG.csv $FULL > $O/traffic_$TAG.log 2>&1
# the pool holds the whole frame: a render is one launch of each of the nine kernels per bounce; -s 9 = the first bounce off a surface
ncu --set full --clock-control none --import-source on -k regex:'k_generate|k_requeue|k_wide_persistent|k_shade' -s 9 -c 9 -o $O/prof_$TAG $SHORT --warmup 0 > $O/ncu_b_$TAG.log 2>&1
$SHORT --mode f32 > $O/plain_${TAG}_f32.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_wide_persistent' -s 2 -c 1 -o $O/prof_${TAG}_f32 $SHORT --warmup 0 --mode f32 > $O/ncu_c_$TAG.log 2>&1
tail -3 $O/traffic_$TAG.log | cut -c1-300; tail -c 600 $O/bench_$TAG.json; echo; cat $O/smoke_$TAG.log | tail -2; tail -2 $O/ncu_b_$TAG.log; tail -2 $O/ncu_c_$TAG.log
