O=gpurun_out
python tools/tune.py base shift3 shift5 blocks4 2>&1 | tee $O/var_linear.log
python tools/tune_lbvh.py 2>&1 | tee $O/var_lbvh.log
for v in fmadefer8; do
  RT_B200_LIB=$PWD/build/variants/librt_b200_$v.so python -m pytest tests -m gpu -x -q -k "lbvh" 2>&1 | tail -3 | tee -a $O/var_lbvh.log
done
