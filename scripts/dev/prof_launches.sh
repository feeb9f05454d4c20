set -u
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r2_a.csv python scripts/run_stage.py all 16 2 > gpurun_out/prof_r2_a.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_r2_a.csv > gpurun_out/launches_r2_a.txt
head -40 gpurun_out/launches_r2_a.txt
