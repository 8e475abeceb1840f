python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke3.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke3.log
python tools/host_time.py 2>&1 | grep "steps:"
