import os, sys, subprocess
# CTA-width sweep: LDPC_B200_THREADS=<cap> python tools/ab/c13.py / one.py
for cap in sys.argv[1:]:
    env = dict(os.environ, LDPC_B200_THREADS=cap)
    print("== cap", cap, flush=True)
    subprocess.run([sys.executable, "tools/ab/c13.py"], env=env)
