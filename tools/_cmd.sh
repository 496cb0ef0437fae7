python tools/kbench.py s6 s3
