set -e
cd /root/repo
for flags in "" "-DB200LP_EXP_NO_STORE" "-DB200LP_EXP_NO_PREFETCH" "-DB200LP_EXP_NO_STORE -DB200LP_EXP_NO_PREFETCH"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -lineinfo -std=c++17 -Xcompiler -fPIC -shared -DB200LP_PHASE_TIMING $flags -o pywr_b200/libb200lp_timing.so pywr_b200/csrc/b200lp.cu
  echo "== flags: $flags"
  python benchmarks/tools/phase_timing.py | tail -9 | tr '\n' ' '; echo
done
