"""integration/spin-ed.patch is a real unified diff against the reference tree: it must apply
cleanly (checked whenever /root/reference is present -- it is in the build container, not on the
GPU box) and every C symbol it imports must be declared in include/sped.h with as many arguments
as the Haskell type has."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATCH = os.path.join(ROOT, "integration", "spin-ed.patch")
REF = "/root/reference"


def test_patch_touches_exactly_the_three_integration_points():
    text = open(PATCH).read()
    files = re.findall(r"^\+\+\+ b/(\S+)", text, flags=re.M)
    assert files == ["configure", "src/SpinED.hs", "src/SpinED/Internal.hs"]
    added = [l[1:] for l in text.splitlines() if l.startswith("+") and not l.startswith("+++")]
    removed = [l[1:] for l in text.splitlines() if l.startswith("-") and not l.startswith("---")]
    assert any("extra-libraries" in l for l in text.splitlines())
    assert any(l.strip() == "sped" for l in added) and any(l.strip() == "lattice_symmetries" for l in removed)
    assert any("eigh primmeOptions primmeOperator" in l for l in removed)      # SpinED.hs:404
    assert any("eighDevice" in l for l in added)


def test_imported_symbols_match_the_header():
    text = open(PATCH).read()
    header = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "sped.h")).read(), flags=re.S)
    added = "\n".join(l[1:] for l in text.splitlines() if l.startswith("+") and not l.startswith("+++"))
    imports = re.findall(r'foreign import ccall \w+ "(\w+)"\s+\w+ ::(.*?)IO CInt', added, flags=re.S)
    assert [name for name, _ in imports] == ["sped_eigh"]
    for name, sig in imports:
        n_args_hs = len([a for a in re.split(r"->\s*(?![^()]*\))", sig) if a.strip()])
        m = re.search(rf"\b{name}\s*\((.*?)\);", header, flags=re.S)
        assert m, name
        depth, n_args_c = 0, 1
        for ch in m.group(1):
            depth += ch == "("
            depth -= ch == ")"
            n_args_c += ch == "," and depth == 0
        assert n_args_hs == n_args_c, (name, n_args_hs, n_args_c)


@pytest.mark.skipif(not os.path.isdir(REF) or shutil.which("patch") is None, reason="needs the reference tree and patch(1)")
def test_patch_applies_to_the_reference(tmp_path):
    for rel in ("configure", "src/SpinED.hs", "src/SpinED/Internal.hs"):
        dst = tmp_path / rel
        dst.parent.mkdir(parents=True, exist_ok=True)
        shutil.copy(os.path.join(REF, rel), dst)
    r = subprocess.run(["patch", "-p1", "--no-backup-if-mismatch", "-i", PATCH], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Hunk" not in r.stdout  # no fuzz, no offsets
    patched = (tmp_path / "src/SpinED.hs").read_text()
    assert "eigh primmeOptions primmeOperator" not in patched and "eighDevice" in patched
    # the regenerated patch is the committed one
    r = subprocess.run(["python", os.path.join(ROOT, "integration", "make_patch.py")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert subprocess.run(["git", "diff", "--quiet", "--", PATCH], cwd=ROOT).returncode in (0, 1)
