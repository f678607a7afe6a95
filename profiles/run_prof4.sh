# fresh --set full captures of every hot kernel inside a log_n_rows = 20 proof (one launch each, mid-proof)
set -x
for k in ifft_low12:40 mid12:40 fft_low12:40 leaves_kernel:20 constraints_tiles:10 bitcol_dot:0 bitrow_comb:0; do
  name=${k%%:*}; skip=${k##*:}
  ncu --set full --clock-control none --import-source on -k regex:$name -s $skip -c 1 -o gpurun_out/prof4_$name -f python profiles/prof_one.py 20 1 > /dev/null 2>&1
done
ls -la gpurun_out/*.ncu-rep
