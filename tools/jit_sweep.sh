QCHECK=1 QMODES= QB=64 python tools/gpu_jit_cfg4.py 2>&1 | grep "steps equal"
run() { echo "$@"; env "$@" QCHECK=0 QMODES=jit QT=20000 python tools/gpu_jit_cfg4.py 2>&1 | tail -1; }
run HY_CUDA_JIT_THREADS=512
run HY_CUDA_JIT_THREADS=256
