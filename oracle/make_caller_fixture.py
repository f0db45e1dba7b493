"""TEST INFRASTRUCTURE -- writes tests/golden/ref_caller_bodies.json.

The drop-in claim of valley_b200/model.py ("swap the import, keep the callers") is only worth something if the
reference's own call sites run against it UNMODIFIED.  Those call sites are Python functions inside scripts that
cannot be imported as modules here or on the GPU box (valley/serve/model_worker.py imports fastapi / decord /
gradio-era helpers at module level; /root/reference does not exist on the GPU box at all).  So, like the golden
tensors, the bodies are extracted once in the build container, by this script, from the sources where they lie:

  valley/serve/model_worker.py   load_model                          (model load + vision tower / token-id setup)
  valley/serve/model_worker.py   ModelWorker.generate_video_stream   (the per-token serving loop)
  valley/inference/run_valley.py init_vision_token, main             (CLI: load, .to(device), .eval(), completion())

tests/test_gpu_dropin.py exec()s each body verbatim in a namespace where ``ValleyLlamaForCausalLM`` is the valley_b200
class and the unrelated third parties (AutoTokenizer, CLIPImageProcessor, decord, logging) are small fakes.  Nothing but
the function text is stored (with file, line range and SHA-256 so drift is visible); it is never imported by the product.

    python oracle/make_caller_fixture.py      # needs /root/reference
"""
import ast
import hashlib
import json
import os
import textwrap

REF = os.environ.get("VALLEY_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = {
    "valley/serve/model_worker.py": ["load_model", "ModelWorker.generate_video_stream"],
    "valley/inference/run_valley.py": ["init_vision_token", "main"],
}


def find(tree, dotted):
    scope = tree.body
    node = None
    for part in dotted.split("."):
        node = next(n for n in scope if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name == part)
        scope = node.body
    return node


def main():
    out = {}
    for rel, names in WANT.items():
        src = open(os.path.join(REF, rel)).read()
        lines = src.splitlines(keepends=True)
        tree = ast.parse(src)
        for name in names:
            node = find(tree, name)
            first = min([node.lineno] + [d.lineno for d in node.decorator_list])
            text = textwrap.dedent("".join(lines[first - 1: node.end_lineno]))
            out[f"{rel}:{name}"] = {"file": rel, "lines": [first, node.end_lineno], "sha256": hashlib.sha256(text.encode()).hexdigest(),
                                    "source": text}
            print(f"{rel}:{first}-{node.end_lineno}  {name}  ({len(text.splitlines())} lines)")
    dst = os.path.join(ROOT, "tests", "golden", "ref_caller_bodies.json")
    json.dump(out, open(dst, "w"), indent=1)
    print("wrote", dst)


if __name__ == "__main__":
    main()
