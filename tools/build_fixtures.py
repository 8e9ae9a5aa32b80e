"""Compile the reference assets into the packaged fixtures the GPU box needs (it has no
/root/reference):  python tools/build_fixtures.py [/root/reference]

  vnl-brax-imitation_b200/data/rodent_model.npz  compiled rodent model (envs/rodent.py:39-63 recipe)
  vnl-brax-imitation_b200/data/rodent_clip.npz   process_clip() of clips/transform_snips_groom.p
  vnl-brax-imitation_b200/data/humanoid_model.npz compiled humanoid model (envs/humanoid.py:40-54 recipe)
  vnl-brax-imitation_b200/data/rodent_pair_model.npz rodent_pair.xml (<replicate count=2>) with the rodent env recipe
  vnl-brax-imitation_b200/data/ant_model.npz     ant.xml, brax-style load (jointless bodies fused; envs/ant.py:40-52 recipe)
  tests/golden/rodent_clip_golden.npz            the old clip's own derived fields (known answers)
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
mjcf = importlib.import_module("vnl-brax-imitation_b200.mjcf")
clipm = importlib.import_module("vnl-brax-imitation_b200.clip")


def main(ref="/root/reference"):
    data = os.path.join(ROOT, "vnl-brax-imitation_b200", "data")
    os.makedirs(data, exist_ok=True)
    xml = os.path.join(ref, "assets", "rodent.xml")
    model = mjcf.load_rodent(xml)
    mjcf.save_model(model, os.path.join(data, "rodent_model.npz"))
    pk = os.path.join(ref, "clips", "transform_snips_groom.p")
    clip = clipm.process_clip(pk, mjcf_path=xml)
    np.savez_compressed(os.path.join(data, "rodent_clip.npz"), **clipm.clip_to_npz_dict(clip))
    old = clipm.load_pickle(pk)
    gold = {k: np.asarray(getattr(old, k)) for k in ("position", "quaternion", "joints", "body_positions", "body_quaternions",
                                                     "center_of_mass", "appendages", "velocity", "angular_velocity",
                                                     "joints_velocity")}
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "rodent_clip_golden.npz"), **gold)
    hum = mjcf.load_humanoid(os.path.join(ref, "assets", "humanoid.xml"))
    mjcf.save_model(hum, os.path.join(data, "humanoid_model.npz"))
    pair = mjcf.load_rodent_pair(os.path.join(ref, "assets", "rodent_pair.xml"))
    mjcf.save_model(pair, os.path.join(data, "rodent_pair_model.npz"))
    ant = mjcf.load_ant(os.path.join(ref, "assets", "ant.xml"))
    mjcf.save_model(ant, os.path.join(data, "ant_model.npz"))
    print("ant", ant.nbody, ant.nv, ant.nu)
    print("model", model.nbody, model.nv, "clip", clip.position.shape, clip.body_positions.shape, "humanoid", hum.nbody, hum.nv)


if __name__ == "__main__":
    main(*sys.argv[1:])
