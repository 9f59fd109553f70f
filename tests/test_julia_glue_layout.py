"""The Julia glue (julia/B200Newton) cannot be executed here (no Julia toolchain): keep at least its struct layouts in
lock-step with include/b200newton.h by comparing field names and order statically."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _c_struct_fields(header, name):
    m = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), header, re.S)
    assert m, name
    body = re.sub(r"/\*.*?\*/", "", m.group(1), flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        names = decl.split(None, 1)[1] if " " in decl else decl
        # "double a, b, c" / "int32_t x" / "b200_gmres_opts gmres"
        typ_and_names = decl.rsplit(" ", 1) if "," not in decl else None
        if typ_and_names:
            fields.append(typ_and_names[1].strip())
        else:
            first, rest = decl.split(",", 1)
            fields.append(first.split()[-1])
            fields += [x.strip() for x in rest.split(",")]
    return [f.lstrip("*") for f in fields]


def _julia_struct_fields(src, name):
    m = re.search(r"struct %s\n(.*?)\nend" % name, src, re.S)
    assert m, name
    return re.findall(r"(\w+)::", m.group(1))


def test_julia_structs_match_the_header():
    header = open(os.path.join(ROOT, "include", "b200newton.h")).read()
    jl = open(os.path.join(ROOT, "julia", "B200Newton", "src", "B200Newton.jl")).read()
    for cname, jname in (("b200_gmres_opts", "GmresOpts"), ("b200_gmres_stats", "GmresStats"), ("b200_newton_opts", "NewtonOpts"),
                         ("b200_newton_result", "NewtonResult")):
        assert _c_struct_fields(header, cname) == _julia_struct_fields(jl, jname), (cname, jname)
    # the positional NewtonOpts(...) call in __solve passes exactly one value per field
    call = jl[jl.index("o = NewtonOpts(something"):]
    call = call[:call.index("alg.alpha_initial)") + len("alg.alpha_initial)")]
    depth, nargs = 0, 1
    for ch in re.sub(r"#.*", "", call[call.index("(") + 1:-1]):
        depth += ch in "(["
        depth -= ch in ")]"
        nargs += ch == "," and depth == 0
    assert nargs == len(_julia_struct_fields(jl, "NewtonOpts"))
