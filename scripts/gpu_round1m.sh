mkdir -p gpurun_out
MBX_LIB=$PWD/metabox_b200/csrc/libmbx_prof.so timeout 600 python scripts/phase_profile.py --fids all,1,21,22 > gpurun_out/phase_r1m.txt 2>&1
cat gpurun_out/phase_r1m.txt
