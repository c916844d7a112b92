QCHECK=1 QMODES= QB=64 python tools/gpu_jit_cfg4.py 2>&1 | grep "steps equal"
for T in 128 256 512; do for C in 1 2; do
  if [ $((T*C)) -le 512 ]; then
    echo "T=$T C=$C"; HY_CUDA_JIT_THREADS=$T HY_CUDA_JIT_CTAS_PER_SM=$C QCHECK=0 QMODES=jit QT=20000 python tools/gpu_jit_cfg4.py 2>&1 | tail -1
  fi
done; done
