#!/usr/bin/env python3
"""Write tests/golden/df_consumer_golden.json: what the REFERENCE'S OWN EstimateAmbientSoundLevel.comp (compiled as C++,
oracle/_ref/libref_shaders.so) returns for fixed listener positions / frames on the stand-in worlds.  Needs /root/reference
(this container only); the JSON travels."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from voxelpathtracer_b200 import assets, world  # noqa: E402
from oracle import ref_shaders  # noqa: E402
from test_df_consumers import listener_positions  # noqa: E402


def main():
    cols = assets.load_plains_columns()
    out = {"source": "Core/Shaders/EstimateAmbientSoundLevel.comp compiled as C++ (oracle/_ref/libref_shaders.so), dispatched as Core/Pipeline.cpp:1921",
           "ambient": {}}
    for name, w in (("gi_box", world.generate_gi_box(cols)), ("city", world.generate_city())):
        df = ref_shaders.df_build(w.data)
        cases = []
        for pos in listener_positions(w, 8, 17) + [(192.0, 75.0, 192.0)]:
            for frame in (0, 5, 640):
                agg, per = ref_shaders.ambient_sound(df, pos, frame)
                cases.append({"pos": list(pos), "frame": frame, "aggregate": agg, "per_invocation": per.tolist()})
        out["ambient"][name] = cases
        print(name, sorted({c["aggregate"] for c in cases}))
    with open(os.path.join(ROOT, "tests", "golden", "df_consumer_golden.json"), "w") as f:
        json.dump(out, f, sort_keys=True)


if __name__ == "__main__":
    main()
