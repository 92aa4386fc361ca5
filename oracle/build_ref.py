"""Recipe for ``oracle/_ref/`` -- the reference's OWN hot-path modules, made runnable.  TEST INFRASTRUCTURE ONLY.

The reference (`/root/reference/bin/scripts/{myCom,myFast5,myDetect}.py`) is Python-2-only source
(``print x``, ``d.has_key(k)``, ``k = d.keys(); k.sort()``) and there is no python2 here.  This
script reads those files WHERE THEY LIE, applies three mechanical py2 -> py3 rewrites, and writes
the result into ``oracle/_ref/`` (git-ignored, not gpurun-ignored: the rendered modules travel to
the GPU box like a built ``.so``; the reference tree itself does not).  Nothing of the reference
is committed.  ``oracle/ref_loader.py`` imports the rendered modules with the scipy-1.2.1 call
semantics injected (``oracle/scipy_legacy.py``) and with stubs for the plotting / HDF5 imports.

Rewrites (nothing else is touched; the test of a faithful rendering is that
``tests/test_oracle_vs_ref.py`` gets identical tables from it and from the restated oracle):
  1. ``print a, b``      -> ``print(a, b)``            (a trailing comma becomes ``end=' '``)
  2. ``obj.has_key(k)``  -> ``(k in obj)``
  3. ``x = d.keys(); x.sort()`` -> ``x = sorted(d.keys())``; any other ``.keys()`` -> ``list(...)``

Run:  python -m oracle.build_ref        (also called by ``__graft_entry__.build()``)
"""
from __future__ import annotations

import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_ref")
REF_ROOT = os.environ.get("NANOMOD_REFERENCE", "/root/reference")
REF_SCRIPTS = os.path.join(REF_ROOT, "bin", "scripts")
FILES = ("myCom.py", "myFast5.py", "myDetect.py")


def _split_top_level(s: str, sep: str) -> int:
    """index of the first ``sep`` of ``s`` outside quotes / brackets, or -1"""
    depth, quote, i = 0, "", 0
    while i < len(s):
        c = s[i]
        if quote:
            if c == "\\":
                i += 1
            elif c == quote:
                quote = ""
        elif c in "'\"":
            quote = c
        elif c == "#":
            return -1
        elif c in "([{":
            depth += 1
        elif c in ")]}":
            depth -= 1
        elif c == sep and depth == 0:
            return i
        i += 1
    return -1


def _find_print(line: str) -> int:
    """index of a py2 print STATEMENT keyword in ``line`` (outside quotes, at statement start or
    after a block colon / semicolon, not followed by a parenthesis), or -1"""
    quote, i, depth = "", 0, 0
    prev = ""  # last non-blank character outside quotes
    while i < len(line):
        c = line[i]
        if quote:
            if c == "\\":
                i += 1
            elif c == quote:
                quote = ""
                prev = c
        elif c in "'\"":
            quote = c
        elif c == "#":
            return -1
        else:
            if c in "([{":
                depth += 1
            elif c in ")]}":
                depth -= 1
            if (depth == 0 and line.startswith("print", i) and prev in ("", ":", ";")
                    and (i == 0 or not (line[i - 1].isalnum() or line[i - 1] == "_"))):
                rest = line[i + 5:]
                if rest[:1] in (" ", "\t") and rest.lstrip()[:1] not in ("(", "=", ""):
                    return i
            if not c.isspace():
                prev = c
        i += 1
    return -1


def _fix_print(line: str) -> str:
    k = _find_print(line)
    if k < 0:
        return line
    pre, rest = line[:k], line[k + 5:].strip()
    tail = ""
    cut = _split_top_level(rest, ";")
    if cut >= 0:
        rest, tail = rest[:cut].rstrip(), rest[cut:]
    hash_at = _split_hash(rest)
    comment = ""
    if hash_at >= 0:
        rest, comment = rest[:hash_at].rstrip(), "  " + rest[hash_at:]
    if rest.endswith(","):
        return "%sprint(%s, end=' ')%s%s" % (pre, rest[:-1].rstrip(), tail, comment)
    return "%sprint(%s)%s%s" % (pre, rest, tail, comment)


def _split_hash(s: str) -> int:
    quote, i = "", 0
    while i < len(s):
        c = s[i]
        if quote:
            if c == "\\":
                i += 1
            elif c == quote:
                quote = ""
        elif c in "'\"":
            quote = c
        elif c == "#":
            return i
        i += 1
    return -1


def _fix_has_key(line: str) -> str:
    while True:
        k = line.find(".has_key(")
        if k < 0:
            return line
        # object expression: walk left over identifiers, dots and balanced brackets
        i, depth = k, 0
        while i > 0:
            c = line[i - 1]
            if c in ")]":
                depth += 1
            elif c in "([":
                if depth == 0:
                    break
                depth -= 1
            elif depth == 0 and not (c.isalnum() or c in "_."):
                break
            i -= 1
        obj = line[i:k]
        # argument: up to the matching parenthesis
        j, depth = k + len(".has_key("), 1
        quote = ""
        while depth:
            c = line[j]
            if quote:
                if c == quote:
                    quote = ""
            elif c in "'\"":
                quote = c
            elif c == "(":
                depth += 1
            elif c == ")":
                depth -= 1
            j += 1
        arg = line[k + len(".has_key("):j - 1]
        line = "%s(%s in %s)%s" % (line[:i], arg, obj, line[j:])


_KEYS_SORT = re.compile(r"(\b\w+) = (.+?)\.keys\(\)\s*;\s*\1\.sort\(\)\s*;?")


def _fix_keys(line: str) -> str:
    line = _KEYS_SORT.sub(lambda m: "%s = sorted(%s.keys())" % (m.group(1), m.group(2)), line)
    if ".keys()" in line and "sorted(" not in line and _split_hash(line.split(".keys()")[0]) < 0:
        # `x = d.keys()` later indexed: py2 returned a list
        line = re.sub(r"= (.+?)\.keys\(\)", lambda m: "= list(%s.keys())" % m.group(1), line, count=1)
    return line


def py2to3(src: str) -> str:
    out = []
    for line in src.split("\n"):
        body = line
        if _split_hash(body.lstrip()) == 0:  # comment line
            out.append(line)
            continue
        body = _fix_keys(body)
        body = _fix_has_key(body)
        body = _fix_print(body)
        out.append(body)
    return "\n".join(out)


def reference_available() -> bool:
    return all(os.path.isfile(os.path.join(REF_SCRIPTS, f)) for f in FILES)


def built() -> bool:
    return all(os.path.isfile(os.path.join(OUT_DIR, f)) for f in FILES)


def build(force: bool = False) -> str | None:
    """Render the reference modules into oracle/_ref/.  Returns the directory, or None when the
    reference tree is absent (the GPU box) and nothing was built earlier."""
    if not reference_available():
        return OUT_DIR if built() else None
    os.makedirs(OUT_DIR, exist_ok=True)
    for f in FILES:
        src_path, dst_path = os.path.join(REF_SCRIPTS, f), os.path.join(OUT_DIR, f)
        if not force and os.path.isfile(dst_path) and os.path.getmtime(dst_path) >= max(
                os.path.getmtime(src_path), os.path.getmtime(__file__)):
            continue
        with open(src_path) as fh:
            text = py2to3(fh.read())
        compile(text, dst_path, "exec")  # a rendering that does not even parse must not be written
        with open(dst_path, "w") as fh:
            fh.write("# GENERATED by oracle/build_ref.py from %s -- do not commit, do not edit\n" % src_path)
            fh.write(text)
    return OUT_DIR


if __name__ == "__main__":
    d = build(force="--force" in sys.argv)
    print("oracle/_ref:", d if d else "reference tree not found at %s" % REF_ROOT)
