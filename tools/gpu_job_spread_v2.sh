# Next-round first measurement: the SRB_SPREAD_V2 loop of the gridding kernel (srb_spread.cuh) against the shipped one.
# Build the variant in the build container first (nvcc cross-compiles):
#   mkdir -p variants && (cd synchrad_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
#       -shared -Xcompiler -fPIC -diag-suppress 177 -DSRB_SPREAD_V2 -o ../../variants/libspread_v2.so srb_api.cu)
# then: gpurun --timeout 200 -- 'bash tools/gpu_job_spread_v2.sh'
mkdir -p gpurun_out
for lib in default variants/libspread_v2.so; do
  if [ "$lib" = default ]; then unset SYNCHRAD_B200_LIB; else export SYNCHRAD_B200_LIB=$PWD/$lib; fi
  python tools/quick_perf.py 592 10000 double spread 2 2>/dev/null | tee -a gpurun_out/spread_v2_perf.txt
done
SYNCHRAD_B200_LIB=$PWD/variants/libspread_v2.so timeout 150 python -m pytest tests/test_gpu_parity.py -q -x -k "gridding or random" 2>&1 | tail -3
