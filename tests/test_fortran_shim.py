"""fortran/mod_moloch_b200.F90 against include/moloch_b200.h, mechanically (no Fortran compiler exists in the
build image): every C entry point has an interface block with the same number of arguments, each of the right
kind and with the VALUE attribute exactly where C passes by value; the bind(C) derived types list the C structs'
fields in order with matching kinds; the enumerators repeat the C enums in order.

Both sides are parsed here independently: the header with a small C-declaration reader, the Fortran source
with the statement preprocessor of the repo's Fortran-subset parser (oracle/refrun/fortran_subset.py)."""
import os
import re

from oracle.refrun.fortran_subset import preprocess, split_top

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
HEADER = os.path.join(ROOT, "include", "moloch_b200.h")
SHIM = os.path.join(ROOT, "fortran", "mod_moloch_b200.F90")


# ---- C side --------------------------------------------------------------------------------------
def parse_header():
    t = re.sub(r"/\*.*?\*/", " ", open(HEADER).read(), flags=re.S)
    t = re.sub(r"#.*", " ", t)
    structs, enums, protos = {}, {}, {}
    for m in re.finditer(r"typedef\s+struct\s*\{(.*?)\}\s*(\w+)\s*;", t, flags=re.S):
        fields = []
        for decl in m.group(1).split(";"):
            decl = " ".join(decl.split())
            if not decl:
                continue
            head = re.match(r"((?:const\s+)?\w+)\s*(.*)", decl)
            for n in head.group(2).split(","):
                n = n.strip()
                fields.append((head.group(1) + "*" * n.count("*"), n.replace("*", "").strip()))
        structs[m.group(2)] = fields
    for m in re.finditer(r"enum\s+(\w+)\s*\{(.*?)\}\s*;", t, flags=re.S):
        enums[m.group(1)] = [e.split("=")[0].strip() for e in m.group(2).split(",") if e.strip()]
    body = re.sub(r"typedef\s+struct\s*\{.*?\}\s*\w+\s*;", " ", t, flags=re.S)
    body = re.sub(r"enum\s+\w+\s*\{.*?\}\s*;", " ", body, flags=re.S)
    body = re.sub(r"typedef[^;]*;", " ", body).replace('extern "C" {', " ")
    for m in re.finditer(r"([\w\s\*]+?)\b(moloch_b200_\w+)\s*\(([^;]*?)\)\s*;", body, flags=re.S):
        ret = " ".join(m.group(1).split()).replace(" *", "*")
        args = []
        a = " ".join(m.group(3).split())
        if a and a != "void":
            for q in a.split(","):
                q = q.strip()
                if "(*" in q:                                   # char (*names)[48]
                    mm = re.match(r"(\w+)\s*\(\*(\w+)\)\[\d+\]", q)
                    args.append((mm.group(1) + "*", mm.group(2)))
                    continue
                mm = re.match(r"(.*?)(\w+)\s*((?:\[\d*\])*)$", q)
                ctype = mm.group(1).strip().replace(" *", "*").replace("* ", "*")
                if mm.group(3):
                    ctype += "*"                                 # T name[n] is a pointer
                args.append((ctype, mm.group(2)))
        protos[m.group(2)] = (ret, args)
    return structs, enums, protos


# ---- Fortran side --------------------------------------------------------------------------------
def parse_shim():
    stmts = preprocess(open(SHIM).read())
    types, enums, funcs = {}, [], {}
    i = 0
    while i < len(stmts):
        s = stmts[i]
        m = re.match(r"type\s*,\s*bind\s*\(\s*c\s*\)\s*::\s*(\w+)", s)
        if m:
            comps = []
            i += 1
            while not stmts[i].startswith("end type"):
                spec, names = stmts[i].split("::")
                for n in split_top(names, ","):
                    comps.append((spec.strip().replace(" ", ""), n.strip()))
                i += 1
            types[m.group(1)] = comps
        elif re.match(r"enum\s*,\s*bind\s*\(\s*c\s*\)", s):
            names = []
            i += 1
            while not stmts[i].startswith("end enum"):
                for n in split_top(stmts[i].split("::")[1], ","):
                    names.append(n.split("=")[0].strip())
                i += 1
            enums.append(names)
        else:
            m = re.match(r"function\s+(moloch_b200_\w+)\s*\((.*?)\)\s*bind\s*\(\s*c\s*,\s*name\s*=\s*'(\w+)'\s*\)\s*result\s*\(\s*(\w+)\s*\)", s)
            if m:
                name, argnames, cname, res = m.group(1), [a.strip() for a in m.group(2).split(",") if a.strip()], m.group(3), m.group(4)
                decl = {}
                i += 1
                while not stmts[i].startswith("end function"):
                    if "::" in stmts[i]:
                        spec, names = stmts[i].split("::")
                        attrs = [a.strip().replace(" ", "") for a in split_top(spec, ",")]
                        for n in split_top(names, ","):
                            n = n.strip()
                            decl[re.sub(r"\(.*\)", "", n)] = (attrs, "(" in n)
                    i += 1
                funcs[name] = (cname, argnames, decl, res)
        i += 1
    return types, enums, funcs


KIND = {"int": "integer(c_int)", "int32_t": "integer(c_int32_t)", "int64_t": "integer(c_int64_t)",
        "uint64_t": "integer(c_int64_t)", "double": "real(c_double)", "char": "character(kind=c_char)"}


def check_arg(fname, ctype, cname, attrs, is_array):
    c = ctype.replace("const", "").replace(" ", "")
    base, by_value = attrs[0], "value" in attrs
    where = f"{fname}({cname}: C `{ctype}`, Fortran `{', '.join(attrs)}`)"
    if not c.endswith("*"):                                    # passed by value in C
        if c == "moloch_b200_physics_fn":
            assert base == "type(c_funptr)" and by_value, where
            return
        assert base == KIND[c] and by_value and not is_array, where
        return
    # a pointer in C: either an opaque address by value, or a by-reference dummy of the pointee's kind
    pointee = c[:-1]
    if base == "type(c_ptr)":
        if pointee.endswith("*"):                              # T** : the address of a pointer
            assert not by_value, where
        else:
            assert by_value, where
        return
    assert not by_value, where
    if pointee in KIND:
        assert base == KIND[pointee], where
    else:
        assert base == f"type({pointee})", where


def test_interfaces_match_the_header():
    structs, enums, protos = parse_header()
    types, fenums, funcs = parse_shim()
    assert len(protos) >= 48
    assert set(funcs) == set(protos), (set(protos) - set(funcs), set(funcs) - set(protos))
    for name, (ret, args) in protos.items():
        cname, argnames, decl, res = funcs[name]
        assert cname == name
        assert len(argnames) == len(args), f"{name}: {len(args)} arguments in C, {len(argnames)} in Fortran"
        for (ctype, ca), fa in zip(args, argnames):
            assert ca.lower() == fa, f"{name}: argument order ({ca} vs {fa})"
            attrs, is_array = decl[fa]
            check_arg(name, ctype, ca, attrs, is_array)
        r = ret.replace("const", "").replace(" ", "")
        want = "type(c_ptr)" if r.endswith("*") else KIND[r]
        assert decl[res][0][0] == want, f"{name}: result {decl[res][0][0]} vs C `{ret}`"


def test_derived_types_mirror_the_structs():
    structs, _, _ = parse_header()
    types, _, _ = parse_shim()
    assert set(structs) == {"moloch_b200_config", "moloch_b200_xfer"} and set(types) == set(structs)
    for name, fields in structs.items():
        comps = types[name]
        assert [f for _, f in fields] == [c for _, c in comps], f"{name}: field order"
        for (ctype, f), (spec, _) in zip(fields, comps):
            want = "type(c_ptr)" if ctype.endswith("*") else KIND[ctype]
            assert spec == want, f"{name}%{f}: {spec} vs C `{ctype}`"


def test_enumerators_repeat_the_enums():
    _, enums, _ = parse_header()
    _, fenums, _ = parse_shim()
    want = [[n.lower() for n in v] for v in enums.values()]
    assert fenums == want
    assert len(want[0]) >= 70 and want[0][-1] == "mb_nfields"      # the full field enum, not a prefix


def test_the_shim_is_complete():
    """No elisions: every helper the module calls is defined in it, every public name exists."""
    src = open(SHIM).read().lower()
    assert "verbatim" not in src and "..." not in src
    for helper in ("put2", "put3", "put4", "get2", "get3", "get4", "upload_state", "download_state", "chk", "c_to_f",
                   "nbr", "l2i", "sp3", "xentry", "putprof"):
        assert re.search(rf"(subroutine|function)\s+{helper}\b", src), helper
    pub = re.findall(r"public\s*::\s*(.*)", src)
    for name in [n.strip() for line in pub for n in line.split(",")]:
        assert re.search(rf"subroutine\s+{name}\b", src), name
    assert max(len(l) for l in open(SHIM).read().splitlines()) <= 132      # free-form line limit
