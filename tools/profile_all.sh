# ncu captures of every hot kernel family (one `--set full` capture per kernel, small launches) and the launch list of the
# headline step.  bash tools/profile_all.sh [outdir]   (on the GPU box; results are summarised with tools/ncu_summary.py)
out=${1:-gpurun_out}
mkdir -p "$out"
cap() {  # name skip count regex mode [count]
  name=$1; skip=$2; cnt=$3; re=$4; shift 4
  timeout 900 ncu --set full --import-source on --clock-control none -k "regex:$re" -s "$skip" -c "$cnt" -o "$out/r02_ncu_$name" -f \
      python tools/prof_kernels.py "$@" > "$out/r02_ncu_$name.log" 2>&1
  echo "$name exit=$?"
  # the reports are tens of MB each: keep the per-launch counters (and, for the headline kernel, the per-instruction
  # sampling page) as CSV and drop the report
  ncu -i "$out/r02_ncu_$name.ncu-rep" --page raw --csv > "$out/r02_ncu_${name}_raw.csv" 2>/dev/null
  case "$name" in enc_fp64|enc_int) ncu -i "$out/r02_ncu_$name.ncu-rep" --page source --csv --print-source sass > "$out/r02_ncu_${name}_sass.csv" 2>/dev/null ;; esac
  rm -f "$out/r02_ncu_$name.ncu-rep"
}
cap enc_fp64 0 1 powm_vm enc
PGPU_NO_FP64=1 cap enc_int 0 1 powm_vm enc
cap dec 1 2 powm_vm dec
cap pdec3072 3 2 powm_vm pdec3072
cap level2 0 6 powm_vm level2 2048
cap ddleq 6 10 powm_vm ddleq 32
cap safeprime 0 3 'strong_kernel|sieve_kernel' safeprime
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$out/r02_ncu_launches.csv" \
    python bench.py --steps 2 --warmup 3 --count 262144 --no-extras > "$out/r02_ncu_launches.log" 2>&1
echo "launch list exit=$?"
