"""The Haskell side of the drop-in boundary, checked as far as it can be without a GHC:

* every `foreign import ccall` of haskell/GPLVM/B200.hs names a function declared in include/gplvm_b200.h and exported by
  the library, with the same arity and matching argument / result types;
* every name in the module's export list is defined in the module;
* haskell/patches/*.patch are what haskell/make_patches.py generates from the reference checkout and apply to it
  (`patch --dry-run`), and the patched sources call the shim (only where /root/reference exists: this container).
"""
import os
import re
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B200 = os.path.join(ROOT, "haskell", "GPLVM", "B200.hs")
HEADER = os.path.join(ROOT, "include", "gplvm_b200.h")
PATCHES = os.path.join(ROOT, "haskell", "patches")
REFERENCE = "/root/reference"

# Haskell FFI type -> C type (ISO C / Haskell 2010 report, section 8.7)
HS2C = {"Ptr Double": {"double*", "const double*"}, "Ptr Int64": {"int64_t*"}, "Ptr CInt": {"int*"}, "Ptr CChar": {"const char*"},
        "Int64": {"int64_t"}, "Double": {"double"}, "CInt": {"int"}}


def c_prototypes():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(const char\*|int64_t|int|void)\s+(gplvm_\w+)\s*\(([^)]*)\)\s*;", text):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        types = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                a = re.sub(r"\s*\b\w+$", "", a) if not a.endswith("*") else a   # drop the parameter name
                a = a.replace(" *", "*").replace("* ", "*")
                types.append(a.strip())
        protos[name] = (ret, types)
    return protos


def hs_imports():
    text = open(B200).read()
    out = {}
    for m in re.finditer(r'foreign import ccall (?:safe|unsafe) "(\w+)"\s+(\w+)\s*::\s*(.*?)(?=\nforeign import|\n\n)', text, flags=re.S):
        cname, hsname, sig = m.group(1), m.group(2), " ".join(m.group(3).split())
        parts = [p.strip() for p in sig.split("->")]
        out[cname] = (hsname, parts)
    return text, out


def test_foreign_imports_match_the_header():
    protos = c_prototypes()
    _, imports = hs_imports()
    assert len(imports) >= 10
    for cname, (hsname, parts) in imports.items():
        assert cname in protos, "%s is not declared in include/gplvm_b200.h" % cname
        ret, ctypes_ = protos[cname]
        *hargs, hret = parts
        assert len(hargs) == len(ctypes_), "%s: arity %d in B200.hs, %d in the header" % (cname, len(hargs), len(ctypes_))
        for i, (h, c) in enumerate(zip(hargs, ctypes_)):
            assert h in HS2C and c in HS2C[h], "%s argument %d: Haskell %r vs C %r" % (cname, i, h, c)
        assert hret.startswith("IO ")
        hr = hret[3:].strip("() ")
        assert hr in HS2C and ret in HS2C[hr], "%s result: Haskell %r vs C %r" % (cname, hret, ret)


def test_foreign_imports_are_exported_by_the_library():
    from gplvmhaskell_b200 import _lib
    lib = _lib.load()
    _, imports = hs_imports()
    for cname in imports:
        getattr(lib, cname)


def test_export_list_is_defined_and_every_import_is_used():
    text, imports = hs_imports()
    exports = re.search(r"module GPLVM\.B200\s*\((.*?)\)\s*where", text, flags=re.S).group(1)
    names = [n.strip() for n in exports.split(",") if n.strip()]
    assert {"ppcaDense", "ppcaMissing", "ppca", "rbfKernel", "gpCholLml", "cholesky", "triSolveLower", "gpPosterior", "pca"} <= set(names)
    for n in names:
        assert re.search(r"^%s\s*::" % re.escape(n), text, flags=re.M), "%s has no type signature" % n
        assert re.search(r"^%s\b[^:\n]*=" % re.escape(n), text, flags=re.M), "%s is exported but not defined" % n
    for cname, (hsname, _) in imports.items():
        assert len(re.findall(r"\b%s\b" % re.escape(hsname), text)) >= 2, "%s is imported but never called" % hsname


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="the reference checkout exists only in the build container")
def test_patches_are_current_and_apply(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "haskell"))
    import make_patches
    generated = make_patches.generate(REFERENCE)
    assert set(generated) == {f for f in os.listdir(PATCHES) if f.endswith(".patch")}
    for name, text in generated.items():
        assert open(os.path.join(PATCHES, name), newline="").read() == text, "%s is stale: run python haskell/make_patches.py" % name
    work = tmp_path / "ref"
    shutil.copytree(REFERENCE, work, ignore=shutil.ignore_patterns(".git"))
    for name in sorted(generated):
        r = subprocess.run(["patch", "-p1", "--dry-run", "-i", os.path.join(PATCHES, name)], cwd=work, capture_output=True, text=True)
        assert r.returncode == 0, (name, r.stdout, r.stderr)
        r = subprocess.run(["patch", "-p1", "-i", os.path.join(PATCHES, name)], cwd=work, capture_output=True, text=True)
        assert r.returncode == 0, (name, r.stdout, r.stderr)
    ppca = (work / "src/GPLVM/PPCA.hs").read_text()
    assert "ppcaDense (computeS learnMatrix)" in ppca and "ppcaMissing (computeS learnMatrix)" in ppca
    assert "makePPCA leaningVectors desiredDimension stopParameter generator" in ppca          # dispatch + RNG draw untouched
    assert "initDispersion = fromIntegral $ fst . next $ generator" in ppca
    typed = (work / "src/GPLVM/TypeSafe/PPCA.hs").read_text()
    for msg in ("dimensions x2 and d1 should be equal, but they are not", "dimensions y1 and y2 should be equal, but they are not",
                "Input matrix should have at least one column"):
        assert msg in typed                                                                       # proofs still run before the FFI
    assert "ppcaMissing (R.computeS $ getInternal learnMatrix)" in typed and "convertPPCATypeSafeData" in typed
    gp = (work / "src/GPLVM/GaussianProcess.hs").read_text()
    assert gp.count("cholGP $") == 2 and gp.count("solveGP cholK") == 2 and "B200.triSolveLower" in gp
    assert "GPLVM.B200" in (work / "GPLVMHaskell.cabal").read_text()
    pca = (work / "src/GPLVM/PCA.hs").read_text()
    assert "B200.pca (computeS input) desiredDimensions" in pca and "= eigSH" not in pca and "= meanCovS" not in pca
    # the emptied bodies: no hmatrix call is left on the EM path of either variant
    for src in (ppca, typed):
        body = src[src.index("emStepsFast"):]
        assert "mulS" not in body and "invS" not in body and "mulM" not in body and "invSM" not in body
