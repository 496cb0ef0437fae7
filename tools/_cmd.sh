timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tools/kbench.py tv:BFMMM_V_CHI=1 tv:BFMMM_V_Z=2 tv:BFMMM_V_SSR=1
