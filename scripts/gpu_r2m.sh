set -x
mkdir -p gpurun_out
timeout 600 python scripts/dropin_timing.py > gpurun_out/r2m_dropin_timing.txt 2>&1; cat gpurun_out/r2m_dropin_timing.txt
timeout 300 python scripts/_debug_gleet_nan.py 2>&1 | tail -5
