#!/usr/bin/env python
"""Prints the ISO_C_BINDING interface blocks, the bind(C) derived types and the enumerators that
include/moloch_b200.h implies (the mechanical part of fortran/mod_moloch_b200.F90).  The file under
fortran/ was assembled from this output; tests/test_fortran_shim.py re-derives both sides independently.

    python scripts/gen_fortran_interfaces.py > /tmp/interfaces.f90
"""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def strip_comments(text):
    return re.sub(r"/\*.*?\*/", " ", text, flags=re.S)


def parse_header(path=os.path.join(ROOT, "include", "moloch_b200.h")):
    """-> (structs {name: [(ctype, field)]}, enums {name: [enumerators]}, protos [(ret, name, [(ctype, arg)])])"""
    t = strip_comments(open(path).read())
    t = re.sub(r"#.*", " ", t)
    structs, enums, protos = {}, {}, []
    for m in re.finditer(r"typedef\s+struct\s*\{(.*?)\}\s*(\w+)\s*;", t, flags=re.S):
        fields = []
        for decl in m.group(1).split(";"):
            decl = " ".join(decl.split())
            if not decl:
                continue
            mm = re.match(r"(.+?)\s*([\w\s,\*]+)$", decl)
            ctype, names = decl.rsplit(" ", 1)[0], None
            # "int32_t a, b, c" / "double* host"
            head = re.match(r"((?:const\s+)?\w+\s*\**)\s*(.*)", decl)
            ctype = head.group(1).replace(" ", "")
            for n in head.group(2).split(","):
                n = n.strip()
                stars = n.count("*")
                fields.append((ctype + "*" * stars, n.replace("*", "").strip()))
        structs[m.group(2)] = fields
    for m in re.finditer(r"enum\s+(\w+)\s*\{(.*?)\}\s*;", t, flags=re.S):
        names = []
        for e in m.group(2).split(","):
            e = e.strip()
            if e:
                names.append(e.split("=")[0].strip())
        enums[m.group(1)] = names
    body = re.sub(r"typedef\s+struct\s*\{.*?\}\s*\w+\s*;", " ", t, flags=re.S)
    body = re.sub(r"enum\s+\w+\s*\{.*?\}\s*;", " ", body, flags=re.S)
    body = re.sub(r"typedef[^;]*;", " ", body)
    body = body.replace('extern "C" {', " ")
    for m in re.finditer(r"([\w\s\*]+?)\b(moloch_b200_\w+)\s*\(([^;]*?)\)\s*;", body, flags=re.S):
        ret = " ".join(m.group(1).split())
        args = []
        a = " ".join(m.group(3).split())
        if a and a != "void":
            for q in a.split(","):
                q = q.strip()
                if "(*" in q:      # char (*names)[48]
                    mm = re.match(r"(\w+)\s*\(\*(\w+)\)\[\d+\]", q)
                    ctype, name, arr = mm.group(1) + "*", mm.group(2), ""
                else:
                    mm = re.match(r"(.*?)(\w+)\s*((?:\[\d*\])*)$", q)
                    ctype, name, arr = mm.group(1).strip(), mm.group(2), mm.group(3)
                ctype = ctype.replace(" *", "*").replace("* ", "*")
                if arr:
                    ctype += "*"
                args.append((ctype, name))
        protos.append((ret, m.group(2), args))
    return structs, enums, protos


KIND = {"int": "integer(c_int)", "int32_t": "integer(c_int32_t)", "int64_t": "integer(c_int64_t)",
        "uint64_t": "integer(c_int64_t)", "double": "real(c_double)"}


def fortran_arg(ctype, name):
    c = ctype.replace("const", "").replace(" ", "")
    if c in KIND:
        return f"{KIND[c]}, value :: {name}"
    if c == "moloch_b200_ctx*" or c == "void*" or c == "double*" and name in ("host",):
        return f"type(c_ptr), value :: {name}"
    if c == "moloch_b200_ctx**":
        return f"type(c_ptr), intent(out) :: {name}"
    if c == "void**":
        return f"type(c_ptr), intent(out) :: {name}"
    if c == "char*":
        return f"character(kind=c_char), intent(in) :: {name}(*)"
    if c == "moloch_b200_physics_fn":
        return f"type(c_funptr), value :: {name}"
    if c.endswith("*") and c[:-1].rstrip("*") in KIND:
        return f"{KIND[c[:-1].rstrip('*')]} :: {name}(*)"
    if c.endswith("*"):
        return f"type({c[:-1]}) :: {name}" + ("(*)" if name in ("down", "up") else "")
    raise ValueError((ctype, name))


def main():
    structs, enums, protos = parse_header()
    for sname, fields in structs.items():
        print(f"  type, bind(C) :: {sname}")
        for ctype, f in fields:
            k = "type(c_ptr)" if ctype.endswith("*") else KIND[ctype]
            print(f"    {k} :: {f}")
        print(f"  end type {sname}\n")
    for ename, names in enums.items():
        print(f"  enum, bind(C)   ! {ename}")
        print(f"    enumerator :: {names[0]} = 0")
        for n in names[1:]:
            print(f"    enumerator :: {n}")
        print("  end enum\n")
    print("  interface")
    for ret, name, args in protos:
        r = ret.replace("const", "").replace(" ", "")
        rk = "type(c_ptr)" if r.endswith("*") else KIND[r]
        line = f"    function {name}({', '.join(a for _, a in args)}) bind(C, name='{name}') result(rc)"
        if len(line) > 120:
            q = line.index(" bind(C")
            line = line[:q] + " &\n        " + line[q + 1:]
        print(line)
        print("      import")
        for ctype, a in args:
            print("      " + fortran_arg(ctype, a))
        print(f"      {rk} :: rc")
        print(f"    end function {name}")
    print("  end interface")


if __name__ == "__main__":
    main()
