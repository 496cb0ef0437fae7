python tools/kbench.py small c5 c6
