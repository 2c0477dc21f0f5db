import os, subprocess, sys
sys.argv = ["x"]
names = {0: "full", 64: "no tmem loads", 128: "no staging stores", 192: "neither", 16: "no bulk store", 17: "no epilogue"}
print(f"{'':28s} 224x512x32  112x256x64  28x64x256  (us)")
for a, n in names.items():
    env = dict(os.environ, QPWC_ABLATE=str(a))
    r = subprocess.run([sys.executable, "tools/ablate_tc.py", "child"], env=env, capture_output=True, text=True, timeout=120)
    print(f"{n:28s} {r.stdout.strip()} {r.stderr.strip()[-200:] if r.returncode else ''}", flush=True)
