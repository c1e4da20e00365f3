"""Developer experiment: time the relaxation kernel under HN_RELAX_VARIANT settings."""
import json, os, subprocess, sys
variants = sys.argv[2:] or ["0"]
probes = sys.argv[1] if len(sys.argv) > 1 else "16384"
for v in variants:
    env = dict(os.environ, HN_RELAX_VARIANT=v)
    r = subprocess.run([sys.executable, "bench.py", "--probes", probes, "--steps", "1", "--warmup", "1", "--no-cpu", "--no-delta", "--no-f32"],
                       capture_output=True, text=True, env=env)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        print("variant", v, "relax_ms", round(d["stage_ms_per_step"]["relax"], 2), "probes/s(relax)",
              round(d["config"]["probes_per_gpu"] / d["stage_ms_per_step"]["relax"] * 1e3), r.stderr.strip().splitlines()[-1:])
    except Exception as e:
        print("variant", v, "failed", e, r.stderr[-500:])
