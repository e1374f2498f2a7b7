# A/B of the exponentiation kernels through bench.py (EncryptWithR + CRT Decrypt, 2^18 items): FP64-pipe shapes against
# the integer-pipe shapes (PGPU_NO_FP64=1).  Run on the GPU box: bash tools/ab_fp64.sh
run() { label=$1; shift; env "$@" python bench.py --count 262144 --steps 2 --warmup 1 --no-extras 2>gpurun_out/ab_err.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$label', round(d['value']), round(d['breakdown']['enc_per_s']), round(d['breakdown']['dec_per_s']))" || tail -5 gpurun_out/ab_err.log; }
run FP64 A=1
[ -n "$AB_ONLY_FP64" ] || run IMAD PGPU_NO_FP64=1
