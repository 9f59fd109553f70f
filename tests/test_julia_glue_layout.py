"""The Julia glue (julia/B200Newton) cannot be executed here (no Julia toolchain).  Static checks that keep it in
lock-step with include/b200newton.h:
  * every mirrored struct has the header's fields in the header's order, with matching scalar types;
  * the option structs are never constructed positionally (only through `with(defaults; field = value ...)`), and every
    keyword passed to `with(...)` names a real field;
  * EVERY `@ccall libb200.<symbol>(...)` names a symbol the header declares, passes exactly as many arguments as the C
    prototype has parameters, with a compatible type in each position, and declares the right return type;
  * the enum tables of the whole-solve algorithm agree with the header's enums.
The executable stand-ins for the glue are the Python harness (same call sequence, GPU tests) and tests/abi_c (plain C)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = open(os.path.join(ROOT, "include", "b200newton.h")).read()
JL = open(os.path.join(ROOT, "julia", "B200Newton", "src", "B200Newton.jl")).read()
HEADER_NC = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)


def _c_struct_fields(name):
    m = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), HEADER_NC, re.S)
    assert m, name
    out = []
    for decl in m.group(1).split(";"):
        decl = " ".join(decl.split())
        if not decl:
            continue
        typ, names = decl.split(" ", 1)
        for nm in names.split(","):
            out.append((nm.strip().lstrip("*"), typ))
    return out


def _julia_struct_fields(name):
    m = re.search(r"\nstruct %s\n(.*?)\nend" % name, JL, re.S)
    assert m, name
    return re.findall(r"(\w+)::(\w+)", m.group(1))


C2J = {"int32_t": "Int32", "int64_t": "Int64", "double": "Float64", "b200_gmres_opts": "GmresOpts"}


def test_julia_structs_match_the_header():
    for cname, jname in (("b200_gmres_opts", "GmresOpts"), ("b200_gmres_stats", "GmresStats"), ("b200_newton_opts", "NewtonOpts"),
                         ("b200_newton_result", "NewtonResult"), ("b200_ens_result", "EnsResult")):
        c = [(n, C2J[t]) for n, t in _c_struct_fields(cname)]
        assert c == _julia_struct_fields(jname), (cname, jname)


def test_option_structs_are_never_built_positionally():
    code = re.sub(r"#.*", "", JL)
    # the only constructor calls allowed: the generic `T(vals...)` inside with(); struct definitions do not count
    for name in ("NewtonOpts", "GmresOpts"):
        for m in re.finditer(r"\b%s\(" % name, code):
            raise AssertionError("positional %s(...) construction at offset %d: use with(default_...(); field = value)" % (name, m.start()))
    fields = {n for n, _ in _julia_struct_fields("NewtonOpts")} | {n for n, _ in _julia_struct_fields("GmresOpts")}
    ncalls = 0
    for m in re.finditer(r"\bwith\(([^;]+);", code):
        depth, i = 1, m.end()
        start = i
        while depth:
            depth += code[i] in "(["
            depth -= code[i] in ")]"
            i += 1
        body = code[start:i - 1]
        if "kw..." in body:
            continue
        ncalls += 1
        d, top = 0, ""
        for ch in body:  # keep only top-level text so nested calls' keywords are not mistaken for fields
            d += ch in "(["
            d -= ch in ")]"
            top += ch if d == 0 else " "
        for kw in re.findall(r"(\w+)\s*=(?!=)", top):
            assert kw in fields, "with(...; %s = ...) names no field of the option structs" % kw
    assert ncalls >= 3


def _split_args(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            depth += ch in "([{"
            depth -= ch in ")]}"
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def _c_prototypes():
    protos = {}
    for m in re.finditer(r"\n\s*(int32_t|void|const char\*|void\*)\s+(b200_\w+)\(([^;{]*?)\);", HEADER_NC, re.S):
        ret, name, params = m.group(1), m.group(2), " ".join(m.group(3).split())
        plist = [] if params in ("", "void") else _split_args(params)
        protos[name] = (ret, plist)
    return protos


def _param_class(p):
    p = p.strip()
    if "*" in p or "_cb " in p or p.endswith("_cb"):
        return "ptr"
    t = p.rsplit(" ", 1)[0].replace("const ", "").strip()
    return {"int32_t": "Int32", "int64_t": "Int64", "double": "Float64", "size_t": "Csize_t"}[t]


def _julia_class(t):
    t = t.strip()
    if t.startswith(("Ptr{", "Ref{")) or t in ("Ctx", "Cstring"):
        return "ptr"
    return t


def test_every_ccall_matches_the_header():
    protos = _c_prototypes()
    assert len(protos) > 90
    seen = set()
    for m in re.finditer(r"@ccall libb200\.(b200_\w+)\(", JL):
        name = m.group(1)
        assert name in protos, "@ccall of %s: not declared in include/b200newton.h" % name
        i, depth = m.end(), 1
        while depth:
            depth += JL[i] in "([{"
            depth -= JL[i] in ")]}"
            i += 1
        args = _split_args(JL[m.end():i - 1])
        ret = re.match(r"::(\w+(\{\w+\})?)", JL[i:]).group(1)
        cret, cparams = protos[name]
        assert len(args) == len(cparams), "%s: %d arguments in the @ccall, %d parameters in the header" % (name, len(args), len(cparams))
        for a, p in zip(args, cparams):
            jt = a.rsplit("::", 1)[1]
            assert _julia_class(jt) == _param_class(p), "%s: argument `%s` vs parameter `%s`" % (name, a, p)
        assert {"int32_t": "Int32", "void": "Cvoid", "const char*": "Cstring", "void*": "Ptr{Cvoid}"}[cret] == ret, (name, ret, cret)
        seen.add(name)
    # the glue binds the whole path: context, vectors, problems, GMRES + preconditioners, Newton, ensemble and its collectives
    for must in ("b200_ctx_create", "b200_gmres_solve", "b200_linop_from_callback", "b200_linop_from_problem", "b200_newton_solve", "b200_newton_solve_host",
                 "b200_ens_solve", "b200_nccl_init", "b200_ens_allgather", "b200_ens_allreduce_stats", "b200_linop_precond", "b200_mul", "b200_extrema"):
        assert must in seen, must


def _c_enum(prefix):
    vals = {}
    for m in re.finditer(r"enum \{([^}]*)\}", HEADER_NC):
        for item in m.group(1).split(","):
            item = item.strip()
            if item.startswith(prefix) and "=" in item:
                k, v = item.split("=")
                vals[k.strip()[len(prefix):].lower()] = int(v)
    return vals


def test_enum_tables_match_the_header():
    def table(name):
        m = re.search(r"const %s = \((.*?)\)\n" % name, JL, re.S)
        return {k: int(v) for k, v in re.findall(r"(\w+) = (\d+)", m.group(1))}
    assert table("_LINSOLVE") == _c_enum("B200_LINSOLVE_")
    assert table("_GLOBALIZATION") == _c_enum("B200_GLOBALIZATION_")
    assert table("_DESCENT") == _c_enum("B200_DESCENT_")
    assert table("_TR_SCHEMES") == _c_enum("B200_TR_")
    assert table("_PRECS") == _c_enum("B200_PRECOND_")
    assert table("_TERMINATION") == _c_enum("B200_TERM_")
    assert table("_QN_INIT") == _c_enum("B200_QN_INIT_")
    assert table("_QN_UPDATE") == _c_enum("B200_QN_UPDATE_")
    # retcode tuple: one entry per B200_RC_* value, in order
    rc = _c_enum("B200_RC_")
    names = re.search(r"const RETCODES = \((.*?)\)\n", JL, re.S).group(1)
    jl_names = re.findall(r"ReturnCode\.(\w+)", names)
    assert len(jl_names) == len(rc) and [n.lower() for n in jl_names] == [k.replace("_", "").replace("linsolve", "linearsolve") for k, _ in sorted(rc.items(), key=lambda kv: kv[1])]
