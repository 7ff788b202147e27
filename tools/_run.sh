PYTHONPATH=. python tools/_dbg3.py 2>&1 | grep -E "walker|spot"
cp lfit_python_b200/liblfit_b200.so /tmp/keep.so; cp lfit_python_b200/liblfit_b200_nowarm.so lfit_python_b200/liblfit_b200.so
echo "--- no warm-up"
PYTHONPATH=. python tools/_dbg3.py 2>&1 | grep -E "spot"
