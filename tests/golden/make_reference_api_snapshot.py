"""Snapshot of the reference's public env API, parsed (ast, nothing imported) from /root/reference -- run in the
BUILD CONTAINER only:   python tests/golden/make_reference_api_snapshot.py  ->  tests/golden/reference_api.json

What is recorded (the drop-in boundary of SURVEY.md section 8(b)): the constructor / reset / step / close / render
signatures of EnhancedRocketTVCEnv (env/enhanced_rocket_tvc_env.py:279-288, :381, :466, :744-753), the MissionPhase
members (:21-29), the names env/__init__.py exports (:105-111), the factory defaults (:66-102) and the three
gym.register ids with their kwargs (:28-64).  tests/test_boundary.py compares the facade with it."""
import ast
import json
import os

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def sig(fn: ast.FunctionDef):
    a = fn.args
    names = [x.arg for x in a.args]
    defaults = [None] * (len(names) - len(a.defaults)) + [ast.unparse(d) for d in a.defaults]
    return [[n, d] for n, d in zip(names, defaults)]


def main():
    env_src = open(os.path.join(REF, "env", "enhanced_rocket_tvc_env.py")).read()
    init_src = open(os.path.join(REF, "env", "__init__.py")).read()
    out = {"source": "env/enhanced_rocket_tvc_env.py + env/__init__.py of NIKHILSAI71/TVC-AI (parsed with ast)"}
    for node in ast.parse(env_src).body:
        if isinstance(node, ast.ClassDef) and node.name == "MissionPhase":
            out["MissionPhase"] = [[t.targets[0].id, ast.literal_eval(t.value)] for t in node.body if isinstance(t, ast.Assign)]
        if isinstance(node, ast.ClassDef) and node.name == "EnhancedRocketTVCEnv":
            out["EnhancedRocketTVCEnv"] = {f.name: sig(f) for f in node.body if isinstance(f, ast.FunctionDef)
                                           and f.name in ("__init__", "reset", "step", "close", "render")}
            out["metadata"] = next(ast.literal_eval(t.value) for t in node.body
                                   if isinstance(t, ast.Assign) and t.targets[0].id == "metadata")
    tree = ast.parse(init_src)
    out["factories"], out["registered"] = {}, {}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name.startswith("make_"):
            call = next(n for n in ast.walk(node) if isinstance(n, ast.Call) and getattr(n.func, "id", "") == "EnhancedRocketTVCEnv")
            defaults = next(ast.literal_eval(n.value) for n in node.body if isinstance(n, ast.Assign) and isinstance(n.value, ast.Dict))
            out["factories"][node.name] = {"signature": sig(node), "vararg_kwargs": node.args.kwarg.arg if node.args.kwarg else None,
                                          "default_kwargs": defaults,
                                          "call_kwargs": {k.arg: ast.unparse(k.value) for k in call.keywords if k.arg}}
        if isinstance(node, ast.Assign) and getattr(node.targets[0], "id", "") == "__all__":
            out["__all__"] = ast.literal_eval(node.value)
    for n in ast.walk(tree):
        if isinstance(n, ast.Call) and getattr(n.func, "id", getattr(n.func, "attr", "")) == "register":
            kw = {k.arg: k.value for k in n.keywords}
            out["registered"][ast.literal_eval(kw["id"])] = {
                "entry_point": ast.literal_eval(kw["entry_point"]), "max_episode_steps": ast.literal_eval(kw["max_episode_steps"]),
                "kwargs": ast.literal_eval(kw["kwargs"])}
    with open(os.path.join(HERE, "reference_api.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(json.dumps(out, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
