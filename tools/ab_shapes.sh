# Per-width A/B of the exponentiation kernels over the rows that are not in bench.py's timed step:
#   bash tools/ab_shapes.sh      (on the GPU box; prints one block per kernel selection)
run() {
  echo "== $*"
  env "$@" python tools/pdec_rate.py 2048 131072 2>&1 | tail -1
  env "$@" python tools/pdec_rate.py 3072 32768 2>&1 | tail -1
  env "$@" python tools/zkp_rate.py 16384 2048 2>&1 | tail -1
  env "$@" python tools/zkp_rate.py 8192 3072 2>&1 | tail -1
  env "$@" python tools/ddleq_rate.py 1024 8 2>&1 | tail -1
}
run A=fp64_everywhere
run PGPU_NO_FP64=1
run PGPU_SHAPE_192=8,24 PGPU_SHAPE_96=4,24
run PGPU_SHAPE_192=16,8,fp64
