"""Documentation consistency (CPU): every profile / tool / source path that DESIGN.md, README.md or INTEGRATION.md cite
exists in the tree, so the evidence a reader is pointed to can actually be opened."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DOCS = ["DESIGN.md", "README.md", "INTEGRATION.md"]
# paths that name the REFERENCE's tree (its configs) or a build product that does not exist for a Python reference
NOT_OURS = {"oracle/_ref", "configs/celebahq/celeb_uncond_ddm_const_uncond_unet_ldm.yaml",
            "configs/super-resolution/div2k_cond_ddm_const_ldm.yaml"}
PATTERN = re.compile(r"`((?:profiles|tools|tests|oracle|adm_b200|include|scripts|configs)/[A-Za-z0-9_./\-]+)`")


def _cited_paths():
    for doc in DOCS:
        text = open(os.path.join(ROOT, doc)).read()
        for m in PATTERN.finditer(text):
            path = m.group(1).rstrip(".")
            if "*" in path or path.endswith("/") or path in NOT_OURS:
                continue
            yield doc, path


def test_cited_paths_exist():
    missing = []
    for doc, path in _cited_paths():
        full = os.path.join(ROOT, path)
        if os.path.exists(full):
            continue
        # `csrc/...` style citations are relative to the package; generated libraries are not in the tree
        if path.endswith(".so") or os.path.exists(os.path.join(ROOT, "adm_b200", path)):
            continue
        missing.append(f"{doc}: {path}")
    assert not missing, "cited but missing:\n" + "\n".join(sorted(set(missing)))
