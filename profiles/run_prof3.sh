ncu --set full --clock-control none --import-source on -k regex:constraints_tiles -s 10 -c 1 -o gpurun_out/prof_cons_r01b python profiles/prof_one.py 18 1 > /dev/null 2>&1
ls -la gpurun_out/prof_cons_r01b.ncu-rep
