mkdir -p gpurun_out
LAYER=conv1 python tools/one_rowseg.py > gpurun_out/g25_run.log 2>&1 || { tail -5 gpurun_out/g25_run.log; exit 1; }
LAYER=conv1 ncu --set full --import-source on --clock-control none -k regex:conv3x3_s1 -s 1 -c 1 -o gpurun_out/g25_conv1 -f python tools/one_rowseg.py > gpurun_out/g25_ncu.log 2>&1
tail -3 gpurun_out/g25_ncu.log
ls -la gpurun_out/
