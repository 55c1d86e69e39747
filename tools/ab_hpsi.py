"""Developer tool: A/B timing of libqsb build variants in ONE process-per-variant loop.

    python tools/ab_hpsi.py N name1=DEF1,DEF2 name2=DEF3 ...   (build here, on the CPU box)
    python tools/ab_hpsi.py --run N name1 name2 ...             (time on the GPU box, interleaved)
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VDIR = os.path.join(ROOT, "qse_b200", "variants")


def child(n):
    import numpy as np

    import qse_b200 as qb
    from qse_b200 import lib

    reg = qb.lattices.square(0.9 * qb.blockade_radius(2 * np.pi), 5, 5) if n == 25 else \
        qb.lattices.random_register(n, 0.9 * qb.blockade_radius(2 * np.pi), seed=n)
    with lib.Engine(n) as eng:
        eng.set_interaction(qb.interaction_matrix(reg.positions))
        if os.environ.get("QSB_SB"):
            eng.set_option("sb_bits", int(os.environ["QSB_SB"]))
        if os.environ.get("QSB_LAG"):
            eng.set_option("sb_lag", int(os.environ["QSB_LAG"]))
        if os.environ.get("QSB_TMAP"):
            eng.set_option("tensor_map", int(os.environ["QSB_TMAP"]))
        if os.environ.get("QSB_TB"):
            eng.set_option("tile_bits", int(os.environ["QSB_TB"]))
        for kv in filter(None, os.environ.get("QSB_OPTS", "").split(",")):  # any qsb_set_option key=value
            eng.set_option(kv.split("=")[0], float(kv.split("=")[1]))
        eng.reset_state()
        eng.evolve_schedule(np.full(2, 3.0), np.full(2, -4.0), np.full(2, 0.001))
        ms = eng.time_apply_h(3.0, -4.0, 14, False)[4:]
        steps = 8
        st = eng.time_schedule(np.linspace(2.0, 6.0, steps), np.linspace(-5, 10, steps), np.full(steps, 0.001))[2:]
        print(json.dumps({"hpsi_ms": round(float(np.median(ms)), 4), "step_ms": round(float(np.median(st)), 3),
                          "terms": eng.metrics()["last_terms"]}))


if __name__ == "__main__":
    if sys.argv[1] == "--child":
        child(int(sys.argv[2]))
    elif sys.argv[1] == "--run":
        n = int(sys.argv[2])
        names = sys.argv[3:]
        for rep in range(3):
            for name in names:
                name, _, opt = name.partition("@")
                opt, _, tb = opt.partition("/")
                sb, _, lag = opt.partition(":")
                if tb:
                    os.environ["QSB_TB"] = tb
                else:
                    os.environ.pop("QSB_TB", None)
                env = dict(os.environ, QSB_LIB=os.path.join(VDIR, f"libqsb_{name}.so"))
                if sb:
                    env["QSB_SB"] = sb
                if lag:
                    env["QSB_LAG"] = lag
                name = f"{name}@{opt}/{tb}"
                out = subprocess.run([sys.executable, __file__, "--child", str(n)], env=env, capture_output=True, text=True)
                print(name, out.stdout.strip() or out.stderr.strip()[-300:], flush=True)
    else:
        from qse_b200 import _build

        os.makedirs(VDIR, exist_ok=True)
        for spec in sys.argv[2:]:
            name, _, defs = spec.partition("=")
            _build.build_variant(os.path.join(VDIR, f"libqsb_{name}.so"), [d for d in defs.split(",") if d])
            print("built", name)
