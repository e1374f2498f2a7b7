# Executed instructions per pipe of the EncryptWithR launch, both multipliers, to reconcile with the program cost
# (SURVEY.md 8d): bash tools/pipe_counts.sh [outdir]
out=${1:-gpurun_out}
M=smsp__inst_executed.sum,smsp__inst_executed_pipe_fmaheavy.sum,smsp__inst_executed_pipe_fma.sum,smsp__inst_executed_pipe_fp64.sum,smsp__inst_executed_pipe_alu.sum,smsp__inst_executed_pipe_lsu.sum,smsp__inst_executed_pipe_xu.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum
ncu --metrics $M --clock-control none -k regex:powm_vm -c 1 --csv --log-file "$out/r02_pipe_counts_fp64.csv" python tools/prof_kernels.py enc > /dev/null 2>&1
PGPU_NO_FP64=1 ncu --metrics $M --clock-control none -k regex:powm_vm -c 1 --csv --log-file "$out/r02_pipe_counts_int.csv" python tools/prof_kernels.py enc > /dev/null 2>&1
grep -h "powm_vm" "$out/r02_pipe_counts_fp64.csv" "$out/r02_pipe_counts_int.csv" | cut -d, -f5,13-15
