#!/usr/bin/env python
"""Generates the `extern "C"` block of bindings/rust/src/ffi.rs from include/zkstark_b200.h, one declaration per
export, so that the Rust view of the ABI cannot drift from the header (tests/test_abi.py checks the two agree).
    python tools/gen_rust_ffi.py            # rewrites the block between the BEGIN/END markers in ffi.rs"""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "zkstark_b200.h")
FFI = os.path.join(ROOT, "bindings", "rust", "src", "ffi.rs")

BASE = {"uint32_t": "u32", "uint64_t": "u64", "uint8_t": "u8", "size_t": "usize", "int": "c_int", "float": "f32", "char": "c_char",
        "void": "c_void", "zk_ctx": "zk_ctx", "zk_shard_comm": "zk_shard_comm", "zk_air": "zk_air"}


def rust_type(ctype, is_array):
    """C parameter type (without the name) -> Rust."""
    t = ctype.strip()
    stars = t.count("*") + (1 if is_array else 0)
    words = t.replace("*", " ").split()
    const_first = words and words[0] == "const"
    base = [w for w in words if w != "const"][0]
    r = BASE[base]
    if stars == 0:
        return r
    # pointer chain, innermost first: `const T *` -> *const T; `T *` -> *mut T; `const void *const *` -> *const *const c_void
    inner_const = const_first
    out = r
    for level in range(stars):
        if level == 0:
            out = ("*const " if inner_const else "*mut ") + out
        else:
            # outer levels: const when the header says `*const` before the next star, or when it is an array of const
            out = ("*const " if "*const" in t.replace(" ", "").replace("const*", "*const") and level == stars - 1 and const_first else "*mut ") + out
    return out


def declarations():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    txt = re.sub(r"^\s*#.*$", "", txt, flags=re.M)
    decls = []
    for m in re.finditer(r"([A-Za-z_][A-Za-z0-9_ \*]*?)\b(zk_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", txt, flags=re.S):
        ret, name, params = m.group(1).strip(), m.group(2), " ".join(m.group(3).split())
        if ret.startswith("typedef"):
            continue
        args = []
        if params and params != "void":
            for p in params.split(","):
                p = p.strip()
                arr = bool(re.search(r"\[[^\]]*\]$", p))
                p = re.sub(r"\[[^\]]*\]$", "", p).strip()
                mm = re.match(r"(.*?)([A-Za-z_][A-Za-z0-9_]*)$", p)
                ctype, pname = mm.group(1), mm.group(2)
                if pname == "type":
                    pname = "type_"
                args.append(f"{pname}: {rust_type(ctype, arr)}")
        if ret == "void":
            rr = ""
        elif ret == "const char *" or ret == "const char*":
            rr = " -> *const c_char"
        else:
            rr = " -> " + rust_type(ret, False)
        decls.append((name, f"    pub fn {name}({', '.join(args)}){rr};"))
    return decls


def main():
    lines = [d for _, d in declarations()]
    block = "    // BEGIN GENERATED (tools/gen_rust_ffi.py from include/zkstark_b200.h)\n" + "\n".join(lines) + "\n    // END GENERATED"
    src = open(FFI).read()
    new = re.sub(r"    // BEGIN GENERATED.*?    // END GENERATED", lambda _: block, src, flags=re.S)
    if "--check" in sys.argv:
        sys.exit(0 if new == src else 1)
    open(FFI, "w").write(new)
    print(f"{len(lines)} declarations")


if __name__ == "__main__":
    main()
