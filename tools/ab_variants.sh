for rep in 1 2; do for a in grid lbvh linear; do python tools/tune.py --accel $a new b0out; done; done
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv
