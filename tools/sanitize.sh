# compute-sanitizer over a small batch of every kernel family, both multiplier paths.
#   bash tools/sanitize.sh [outdir]      (on the GPU box)
# memcheck: out-of-bounds / misaligned accesses anywhere; racecheck: the shared-memory exchange of the dedicated squaring
# (Mont::sqr, __syncwarp fences) and of prod_reduce_kernel.
out=${1:-gpurun_out}
mkdir -p "$out"
run() {  # tool label env... -- mode count
  tool=$1; label=$2; shift 2
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  echo "== $tool $label ${envs[*]} : prof_kernels.py $*" | tee -a "$out/r02_sanitizer_summary.txt"
  env "${envs[@]}" timeout 900 compute-sanitizer --tool "$tool" --print-limit 5 python tools/prof_kernels.py "$@" > "$out/r02_sanitizer_${tool}_${label}.log" 2>&1
  echo "exit=$?" >> "$out/r02_sanitizer_${tool}_${label}.log"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|^ok |exit=" "$out/r02_sanitizer_${tool}_${label}.log" | tee -a "$out/r02_sanitizer_summary.txt"
}
: > "$out/r02_sanitizer_summary.txt"
run memcheck enc_fp64 A=1 -- enc 300
run memcheck enc_int PGPU_NO_FP64=1 -- enc 300
run memcheck enc_int_nosqr PGPU_NO_FP64=1 PGPU_NO_SQR=1 -- enc 300
run memcheck all_fp64 PGPU_SHAPE_64=4,10,fp64 PGPU_SHAPE_96=4,15,fp64 PGPU_SHAPE_192=8,15,fp64 -- level2 64
run memcheck level2 A=1 -- level2 64
run memcheck ddleq A=1 -- ddleq 3
run memcheck zkp3072 A=1 -- zkp 24
run memcheck pdec3072 A=1 -- pdec3072 40
run memcheck safeprime A=1 -- safeprime 512
run memcheck light A=1 -- light 700
run racecheck dec_sqr A=1 -- dec 300
run racecheck level2 A=1 -- level2 48
run racecheck light A=1 -- light 300
