#!/bin/bash
# closing evidence (gpurun_out must stay under 64 MiB: ncu reports are condensed to CSV on the box and deleted)
TAG=${1:-v16}
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --cfg3-frames 0 > gpurun_out/r2_${TAG}_bench.json 2> gpurun_out/r2_${TAG}_bench.err
echo "bench rc=$?"; tail -2 gpurun_out/r2_${TAG}_bench.err
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_${TAG}_launches.csv python scripts/prof_all.py > gpurun_out/r2_${TAG}_ncu1.log 2>&1
echo "launch list rc=$?"; tail -3 gpurun_out/r2_${TAG}_ncu1.log
timeout 420 ncu --set full --clock-control none -c 26 -k regex:'mbw_warp|mbs_decide|mbs_warp|mbs_lap|pyrdown0_tma|mbs_mark|mbs_pull|rnd_warp|rnd_blend|rnd_final|rnd_pyrdown' -o /tmp/r2_${TAG}_full python scripts/prof_all.py > gpurun_out/r2_${TAG}_ncu2.log 2>&1
echo "full rc=$?"; tail -3 gpurun_out/r2_${TAG}_ncu2.log
python scripts/ncu_summary.py /tmp/r2_${TAG}_full.ncu-rep > gpurun_out/r2_${TAG}_ncu_full.csv 2> gpurun_out/r2_${TAG}_ncu_summary.err
ls -la /tmp/r2_${TAG}_full.ncu-rep gpurun_out/ | tail -12
s=$(stat -c %s /tmp/r2_${TAG}_full.ncu-rep 2>/dev/null || echo 0)
if [ "$s" -gt 0 ] && [ "$s" -lt 40000000 ]; then cp /tmp/r2_${TAG}_full.ncu-rep gpurun_out/; fi
du -sh gpurun_out
