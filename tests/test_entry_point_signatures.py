"""pygcransac's entry points against the reference's pybind11 declarations (src/pygcransac/src/bindings.cpp): the same
argument names, in the same order, with the same defaults; anything this package adds (seed, device, scoring, ...) comes
after them, with a default.  CPU tier; only where the reference tree exists."""
import inspect
import os
import re

import pytest

BINDINGS = os.path.join(os.environ.get("GCR_REFERENCE_DIR", "/root/reference"), "src", "pygcransac", "src", "bindings.cpp")
pytestmark = pytest.mark.skipif(not os.path.exists(BINDINGS), reason="no reference tree")


def _reference_arguments(name):
    text = open(BINDINGS).read()
    block = re.search(r'm\.def\("%s",(.*?)\);' % name, text, re.S).group(1)
    return re.findall(r'py::arg\("(\w+)"\)\s*(?:=\s*([^,\n]+?))?\s*(?:,|\n|$)', block)


def _value(cpp_literal):
    return {"true": True, "false": False}.get(cpp_literal, None) if cpp_literal in ("true", "false") else float(cpp_literal)


@pytest.mark.parametrize("name", ["findHomography", "findFundamentalMatrix", "findEssentialMatrix", "find6DPose"])
def test_signature_matches_bindings(name):
    import pygcransac
    ref = _reference_arguments(name)
    ours = list(inspect.signature(getattr(pygcransac, name)).parameters.values())
    assert len(ref) >= 14 and len(ours) >= len(ref)
    for (arg, default), p in zip(ref, ours):
        assert p.name == arg, (name, arg, p.name)
        if default:
            assert p.default is not inspect.Parameter.empty and p.default == _value(default), (name, arg, default, p.default)
        elif arg != "probabilities":            # (the reference takes `probabilities` positionally; an empty default is a convenience)
            assert p.default is inspect.Parameter.empty, (name, arg)
    for p in ours[len(ref):]:
        assert p.default is not inspect.Parameter.empty, (name, p.name)     # additions never change a reference call


def _cpp_parameters(path, name):
    text = open(path).read()
    text = re.sub(r"//[^\n]*", "", text)
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    block = re.search(r"\bint\s+%s\s*\((.*?)\)\s*(?:;|\{)" % name, text, re.S).group(1)
    out = []
    for p in block.split(","):
        p = " ".join(p.split())
        if not p:
            continue
        default = None
        if "=" in p:
            p, default = [x.strip() for x in p.split("=")]
        words = p.replace("&", " & ").split()
        out.append((" ".join(w for w in words[:-1] if w != "const").replace(" &", "&"), words[-1], default))
    return out


@pytest.mark.parametrize("name", ["findHomography_", "findFundamentalMatrix_", "findEssentialMatrix_", "find6DPose_"])
def test_cpp_entry_point_matches_gcransac_python_h(name):
    """host/gcr_gcransac_python.h against the reference's src/pygcransac/include/gcransac_python.h: parameter types and names
    in the reference's order, then `seed` and `device`, defaulted (a C++ caller switches headers and relinks)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ref = _cpp_parameters(os.path.join(os.path.dirname(os.path.dirname(BINDINGS)), "include", "gcransac_python.h"), name)
    ours = _cpp_parameters(os.path.join(root, "graph-cut-ransac_b200", "host", "gcr_gcransac_python.h"), name)
    assert len(ref) >= 16 and [(t, n) for t, n, _ in ours[: len(ref)]] == [(t, n) for t, n, _ in ref], (name, ref, ours)
    assert [(n, d is not None) for _, n, d in ours[len(ref):]] == [("seed", True), ("device", True)], (name, ours[len(ref):])
