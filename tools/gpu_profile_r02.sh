#!/usr/bin/env bash
# GPU box: ncu evidence for round 2 -> gpurun_out/ (summaries are copied into profiles/ afterwards)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/profile_run.py bench 2 > /dev/null 2>&1   # generate + cache the trees outside the profiler
python tools/profile_run.py tt 2 > /dev/null 2>&1
# launch list of one bench.py step sequence (per-launch durations; shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 12 --warmup 3 --no-baselines --no-extras --no-cli --min-seconds 0.01 > gpurun_out/ncu_bench.log 2>&1
for cfg in bench spp1 tt; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 6 -c 2 -f -o gpurun_out/r02_render_$cfg \
      python tools/profile_run.py $cfg 10 > gpurun_out/ncu_render_$cfg.log 2>&1
  ncu -i gpurun_out/r02_render_$cfg.ncu-rep --page raw --csv > gpurun_out/r02_render_$cfg.raw.csv 2>/dev/null
  python tools/ncu_summary.py gpurun_out/r02_render_$cfg.raw.csv > gpurun_out/r02_render_${cfg}_ncu_full.txt
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"guidance_net|filter_sep" -s 12 -c 4 -f -o gpurun_out/r02_denoise \
    python tools/profile_run.py bench 10 > gpurun_out/ncu_denoise.log 2>&1
ncu -i gpurun_out/r02_denoise.ncu-rep --page raw --csv > gpurun_out/r02_denoise.raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r02_denoise.raw.csv > gpurun_out/r02_denoise_ncu_full.txt
timeout 600 ncu --set full --clock-control none -k regex:"guidance_net|filter_sep" -s 12 -c 2 -f -o gpurun_out/r02_denoise_tt \
    python tools/profile_run.py tt 10 > gpurun_out/ncu_denoise_tt.log 2>&1
ncu -i gpurun_out/r02_denoise_tt.ncu-rep --page raw --csv > gpurun_out/r02_denoise_tt.raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r02_denoise_tt.raw.csv > gpurun_out/r02_denoise_tt_ncu_full.txt
rm -f gpurun_out/*.raw.csv gpurun_out/r02_render_spp1.ncu-rep gpurun_out/r02_render_tt.ncu-rep gpurun_out/r02_denoise_tt.ncu-rep   # gpurun_out is capped at 64 MiB
ls -la gpurun_out | head -40
head -30 gpurun_out/r02_render_bench_ncu_full.txt
