# end-of-round --set full captures (log_n_rows = 20): the kernels that changed after run_prof4.sh, plus pass C
set -x
for k in '^fft_low12_kernel:40' 'mid12_kernel:40' 'constraints_tiles:10' 'bitrow_lookup:0'; do
  name=${k%%:*}; skip=${k##*:}; tag=$(echo $name | tr -d '^')
  ncu --set full --clock-control none --import-source on -k "regex:$name" -s $skip -c 1 -o gpurun_out/prof5_$tag -f python profiles/prof_one.py 20 1 > /dev/null 2>&1
done
ls -la gpurun_out/prof5_*.ncu-rep
