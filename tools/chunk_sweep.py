"""Measurement aid: bench.py at several tensor-core chunk sizes (UFO_TC_CHUNK), one line each."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for c in [int(x) for x in sys.argv[1:]] or [4736, 9472, 14208, 18944]:
    env = dict(os.environ, UFO_TC_CHUNK=str(c))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--no-cpu-baseline", "--no-costvolume", "--no-accuracy"],
                       env=env, capture_output=True, text=True)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        print(f"chunk {c}: {d['value']:.4g} rays/s, {d['ms_per_step']:.1f} ms/step, e2e {d['e2e']['value']:.4g}", flush=True)
    except Exception as ex:  # noqa: BLE001
        print(f"chunk {c}: failed ({ex}) {r.stderr[-300:]}", flush=True)
