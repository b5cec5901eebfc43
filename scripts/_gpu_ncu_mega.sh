mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_mega -s 20 -c 1 -f -o gpurun_out/ncu_mega python scripts/ncu_mega.py > gpurun_out/ncu_mega.log 2>&1
tail -3 gpurun_out/ncu_mega.log
ls -la gpurun_out/ncu_mega.ncu-rep
