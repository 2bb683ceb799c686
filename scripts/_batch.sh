set -x
mkdir -p gpurun_out/r2b
python scripts/c5_debug.py c5 1000 2500 > gpurun_out/r2b/c5_debug.json 2> gpurun_out/r2b/c5_debug.err
python scripts/c5_debug.py c3 4000 > gpurun_out/r2b/c3_debug.json 2> gpurun_out/r2b/c3_debug.err
python scripts/gpu_parity_report.py > gpurun_out/r2b/parity_default.log 2>&1
FWI_B200_LIB=$PWD/variants/libfwi_f64near8.so python scripts/gpu_parity_report.py > gpurun_out/r2b/parity_f64near8.log 2>&1
FWI_B200_LIB=$PWD/variants/libfwi_f64near8.so python scripts/parity_full_length.py marmousi 8 > gpurun_out/r2b/marmousi_f64near8.json 2> gpurun_out/r2b/marmousi_f64near8.err
python scripts/parity_full_length.py marmousi 8 > gpurun_out/r2b/marmousi8.json 2> gpurun_out/r2b/marmousi8.err
cat gpurun_out/r2b/*.json | cut -c1-1500
grep -E "^\[|grad_stf" gpurun_out/r2b/parity_default.log gpurun_out/r2b/parity_f64near8.log
