for k in 1 2 4; do echo "==== KBS=$k"; DFU_G2_KBS=$k timeout 200 python scripts/exp_pair.py 2>&1 | grep -E "conv|linear|, 0, 2\)|, 1\) " ; done
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
