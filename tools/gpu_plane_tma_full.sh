# ncu --set full of the TMA-staged Tier-2 leapfrog (one launch after warm-up) + the kinetic golden test; see profiles/README.md
export PYTHONDONTWRITEBYTECODE=1
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_golden.py -q -m gpu -k kinetic 2>&1 | tail -3
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_leapfrog_tma --launch-skip 3 -c 1 -f -o gpurun_out/prof_plane_tma_r5c python tools/plane_leapfrog_run.py 2>&1 | tail -2
