"""TEST INFRASTRUCTURE ONLY.  Rewrites the CUDA launch syntax of a .cu file into calls of the emulation runtime
(tests/emu/cuda_emu.h):

    name<<<grid, block, smem, stream>>>(args);   ->   emu::run_grid_cfg(emu::Cfg(grid, block, smem, stream), [&]() { name(args); });

and the file-scope `extern __shared__ T name[];` declarations into accessors of the block's dynamic shared memory.
usage: transform.py in.cu out.cpp
"""
import re
import sys


def match_close(s, i, open_ch, close_ch):
    """s[i] == open_ch; index of the matching close_ch (skipping strings / chars)"""
    depth = 0
    k = i
    while k < len(s):
        c = s[k]
        if c in "\"'":
            q = c
            k += 1
            while s[k] != q:
                if s[k] == "\\":
                    k += 1
                k += 1
        elif c == open_ch:
            depth += 1
        elif c == close_ch:
            depth -= 1
            if depth == 0:
                return k
        k += 1
    raise ValueError("unbalanced " + open_ch)



def split_top(text):
    """split at commas that are not nested in (), <>, [] or {}"""
    parts, depth, cur = [], 0, []
    for ch in text:
        if ch in "([{<":
            depth += 1
        elif ch in ")]}>":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append("".join(cur)); cur = []
        else:
            cur.append(ch)
    parts.append("".join(cur))
    return parts

def name_start(s, i):
    """start of the kernel name that ends right before s[i:] == '<<<'"""
    k = i
    if s[k - 1] == ">":                     # template arguments
        depth = 0
        k -= 1
        while True:
            if s[k] == ">":
                depth += 1
            elif s[k] == "<":
                depth -= 1
                if depth == 0:
                    break
            k -= 1
    while k > 0 and (s[k - 1].isalnum() or s[k - 1] in "_:"):
        k -= 1
    return k


FILE_ID = 0


def transform(src):
    out = []
    pos = 0
    while True:
        i = src.find("<<<", pos)
        if i < 0:
            out.append(src[pos:])
            break
        ns = name_start(src, i)
        j = src.index(">>>", i)
        cfg = src[i + 3:j]
        p = j + 3
        while src[p].isspace():
            p += 1
        assert src[p] == "(", src[i - 40:i + 80]
        q = match_close(src, p, "(", ")")
        name, args = src[ns:i], src[p + 1:q]
        out.append(src[pos:ns])
        # a launch that asks for dynamic shared memory also hands over the kernel's address, so that the emulator can
        # check the opt-in (cudaFuncSetAttribute) the hardware insists on above 48 KB
        parts = split_top(cfg)
        if len(parts) >= 3 and parts[2].strip() not in ("0", ""):
            out.append("emu::run_grid_cfg(emu::Cfg(%s), [&]() { %s(%s); }, emu::fn_key(%s))" % (cfg, name, args, name))
        else:
            out.append("emu::run_grid_cfg(emu::Cfg(%s), [&]() { %s(%s); })" % (cfg, name, args))
        pos = q + 1
    s = "".join(out)
    # dynamic shared memory: file scope -> accessor macro, inside a function -> local pointer
    def dyn(m):
        indent, typ, name = m.group(1), m.group(2).strip(), m.group(3)
        if indent == "":
            return "#define %s ((%s*)emu::dyn_smem())" % (name, typ)
        return "%s%s* const %s = (%s*)emu::dyn_smem();" % (indent, typ, name, typ)
    s = re.sub(r"(?m)^([ \t]*)extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?(\w[\w \t\*]*?)\s+(\w+)\[\];", dyn, s)
    # static shared memory: one per-block slot per declarator
    site = [0]
    def stat(m):
        indent, typ, decls = m.group(1), m.group(2).strip(), m.group(3)
        outl = []
        for d in [x.strip() for x in decls.split(",")]:
            mm = re.match(r"(\w+)((?:\[[^\]]*\])*)$", d)
            assert mm, d
            name, dims = mm.group(1), mm.group(2)
            site[0] += 1
            sid = FILE_ID * 1000 + site[0]
            full = "%s%s" % (typ, dims)
            if indent == "":
                assert dims == "", "file-scope __shared__ arrays are not handled: " + d
                outl.append("#define %s (*reinterpret_cast<%s*>(emu::shared_slot(%d, sizeof(%s))))" % (name, typ, sid, typ))
            elif dims:
                outl.append("%s%s (&%s)%s = *reinterpret_cast<%s (*)%s>(emu::shared_slot(%d, sizeof(%s)));" % (indent, typ, name, dims, typ, dims, sid, full))
            else:
                outl.append("%s%s& %s = *reinterpret_cast<%s*>(emu::shared_slot(%d, sizeof(%s)));" % (indent, typ, name, typ, sid, typ))
        return "\n".join(outl)
    s = re.sub(r"(?m)^([ \t]*)__shared__\s+(?:__align__\(\d+\)\s+)?(\w[\w \t]*?\**)\s+(\w+(?:\[[^\]]*\])*(?:\s*,\s*\w+(?:\[[^\]]*\])*)*);", stat, s)
    assert "__shared__" not in re.sub(r"//.*", "", s), [l for l in s.splitlines() if "__shared__" in l and not l.strip().startswith("//")][:3]
    s = s.replace('"../../include/clarabel_b200.h"', '"clarabel_b200.h"')      # gen/ sits elsewhere: -Iinclude
    return s


if __name__ == "__main__":
    import zlib
    FILE_ID = zlib.crc32(sys.argv[1].encode()) % 2000 + 1
    src = open(sys.argv[1]).read()
    open(sys.argv[2], "w").write("// GENERATED by tests/emu/transform.py from %s -- do not edit\n" % sys.argv[1] + transform(src))
