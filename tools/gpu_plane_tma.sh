export PYTHONDONTWRITEBYTECODE=1
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_primitives.py -q -m gpu -k "leapfrog" 2>&1 | tail -15
for f in 1 0; do
  NUTS_B200_PLANE_TMA=$f timeout 150 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_leapfrog --launch-skip 3 -c 3 --csv --log-file gpurun_out/plane_leapfrog_tma${f}_r5a.csv python tools/plane_leapfrog_run.py 2>&1 | tail -2
  grep -E "gpu__time_duration" gpurun_out/plane_leapfrog_tma${f}_r5a.csv | awk -F'","' '{print $5, $NF}' | cut -c1-60,200-
done
