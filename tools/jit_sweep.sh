run() { echo "$@"; env "$@" QCHECK=0 QMODES=jit QB=1000000 QT=6000 python tools/gpu_jit_cfg4.py 2>&1 | tail -1; }
run A=1
run HY_CUDA_JIT_BLOCK=4
run HY_CUDA_JIT_BLOCK=2
