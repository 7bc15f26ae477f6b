#!/usr/bin/env python
"""Extracts the signatures of the reference's Task plugin hooks from /root/reference/train.py with `ast` (no import: ksim /
jax are not installable here) and writes tests/golden/ref_signatures.json -- the fixture tests/test_host_cpu.py compares
kbot_joystick_b200.ksim_adapter and kbot_joystick_b200.task against.  Re-run when the reference changes."""
import ast
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
REF = Path(sys.argv[1]) if len(sys.argv) > 1 else Path("/root/reference/train.py")
HOOKS = ("get_observations", "get_commands", "get_rewards", "get_terminations", "get_actuators", "get_model",
         "get_initial_model_carry", "sample_action", "get_ppo_variables", "run_actor", "run_critic", "_ppo_scan_fn",
         "normalize_joint_pos", "normalize_joint_vel", "encode_projected_gravity", "mirror_joints", "mirror_obs", "mirror_cmd",
         "get_optimizer")


def main():
    tree = ast.parse(REF.read_text())
    out = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.ClassDef) and node.name == "HumanoidWalkingTask":
            for fn in node.body:
                if isinstance(fn, ast.FunctionDef) and fn.name in HOOKS:
                    a = fn.args
                    out[fn.name] = {"args": [x.arg for x in a.posonlyargs + a.args], "kwonly": [x.arg for x in a.kwonlyargs],
                                    "lineno": fn.lineno, "returns": ast.unparse(fn.returns) if fn.returns else None}
    dst = ROOT / "tests" / "golden" / "ref_signatures.json"
    dst.write_text(json.dumps({"source": "kscalelabs/kbot-joystick train.py (class HumanoidWalkingTask)", "hooks": out}, indent=1) + "\n")
    print(f"{len(out)} hooks -> {dst}")


if __name__ == "__main__":
    main()
