#!/usr/bin/env python3
"""Decode the reference's bundled .rda fixtures into neutral JSON under tests/golden/.

Run ONCE in the build container (needs /root/reference, which does not exist on the GPU box):

    python tools/rda_to_golden.py [/root/reference] [tests/golden]

The four files (`data/NPannual.rda`, `data/NPpc.rda`, `data/theta.rda`, `R/sysdata.rda`) are
bzip2-compressed R serialization v2 in XDR form ("RDX2\\nX\\n").  R is not installed here, so
this is a ~100-line pure-Python reader of exactly the SEXP types those files contain.
Floats are emitted with repr() (round-trip exact, 17 significant digits).

Outputs (all committed):
  tests/golden/np.json      NPannual (year, Qa), NPpc (PC1, PC9, PC13), example theta
  tests/golden/p1.json      P1annual (year, Qa), P1pc (7 PCs x 406)   [R/sysdata.rda]
  tests/golden/nplds.json   NPlds: precomputed LDS_reconstruction result (rec, theta, lik)
  tests/golden/npcv.json    NPcv: precomputed cvLDS result (30 folds Z, Ycv, metrics)
  tests/golden/kat.json     the known-answer numbers of tests/testthat/test-LDS-EM.R:26-40
"""
import bz2
import json
import os
import struct
import sys


class _Reader:
    def __init__(self, raw):
        self.b = raw
        self.o = 0
        self.refs = []

    def i32(self):
        v = struct.unpack_from(">i", self.b, self.o)[0]
        self.o += 4
        return v

    def f64s(self, n):
        v = struct.unpack_from(">%dd" % n, self.b, self.o)
        self.o += 8 * n
        return list(v)

    def i32s(self, n):
        v = struct.unpack_from(">%di" % n, self.b, self.o)
        self.o += 4 * n
        return list(v)

    def item(self):
        flags = self.i32()
        ty = flags & 0xFF
        has_attr = bool(flags & (1 << 9))
        has_tag = bool(flags & (1 << 10))
        if ty == 254:  # NILVALUE
            return None
        if ty in (253, 242):  # global / empty env
            return {"_env": ty}
        if ty == 255:  # REFSXP
            idx = flags >> 8
            if idx == 0:
                idx = self.i32()
            return self.refs[idx - 1]
        if ty == 1:  # SYMSXP
            name = self.item()
            sym = {"_sym": name}
            self.refs.append(sym)
            return sym
        if ty == 2:  # LISTSXP (pairlist): iterate nodes until NILVALUE
            out = []
            while True:
                attr = self.item() if has_attr else None  # noqa: F841 (attributes on nodes unused)
                tag = self.item() if has_tag else None
                car = self.item()
                out.append((tag["_sym"] if tag else None, car))
                flags = self.i32()
                ty = flags & 0xFF
                if ty == 254:
                    break
                if ty != 2:
                    raise ValueError("pairlist cdr of type %d" % ty)
                has_attr = bool(flags & (1 << 9))
                has_tag = bool(flags & (1 << 10))
            return {"_pairlist": out}
        if ty == 9:  # CHARSXP
            n = self.i32()
            if n == -1:
                return None
            s = self.b[self.o:self.o + n].decode("utf-8")
            self.o += n
            return s
        if ty in (10, 13):  # LGLSXP / INTSXP
            val = self.i32s(self.i32())
        elif ty == 14:  # REALSXP
            val = self.f64s(self.i32())
        elif ty == 16:  # STRSXP
            val = [self.item() for _ in range(self.i32())]
        elif ty == 19:  # VECSXP
            val = [self.item() for _ in range(self.i32())]
        elif ty == 22:  # EXTPTRSXP (data.table .internal.selfref)
            ext = {"_extptr": True}
            self.refs.append(ext)
            self.item()
            self.item()
            val = ext
        else:
            raise ValueError("unsupported SEXP type %d at offset %d" % (ty, self.o))
        attrs = {}
        if has_attr:
            pl = self.item()
            if pl is not None:
                attrs = {k: v for k, v in pl["_pairlist"]}
        return _simplify(val, attrs)


def _simplify(val, attrs):
    """Named lists -> dict; matrices -> {'dim':..., 'data': column-major}; else plain list."""
    if isinstance(val, dict):
        return val
    names = attrs.get("names")
    if names is not None and len(names) == len(val) and all(isinstance(n, str) for n in names):
        d = dict(zip(names, val))
        d.pop(".internal.selfref", None)
        return d
    if "dim" in attrs:
        out = {"dim": attrs["dim"], "data": val}
        dn = attrs.get("dimnames")
        if dn is not None:
            out["dimnames"] = dn
        return out
    return val


def load_rda(path):
    raw = bz2.decompress(open(path, "rb").read())
    if raw[:5] != b"RDX2\n" or raw[5:7] != b"X\n":
        raise ValueError("%s: not an RDX2/XDR file" % path)
    r = _Reader(raw)
    r.o = 7
    r.i32()
    r.i32()
    r.i32()  # serialization version, writer R version, min reader version
    top = r.item()
    return {k: v for k, v in top["_pairlist"]}


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(__file__), "..", "tests", "golden")
    os.makedirs(out, exist_ok=True)
    np_annual = load_rda(os.path.join(ref, "data/NPannual.rda"))["NPannual"]
    np_pc = load_rda(os.path.join(ref, "data/NPpc.rda"))["NPpc"]
    theta = load_rda(os.path.join(ref, "data/theta.rda"))["theta"]
    sysd = load_rda(os.path.join(ref, "R/sysdata.rda"))

    def mat1(m):  # 1xN R matrix or plain vector -> list
        return m["data"] if isinstance(m, dict) and "data" in m else m

    def theta_json(th):
        return {k: mat1(th[k]) for k in ("A", "B", "C", "D", "Q", "R", "mu1", "V1")}

    def dump(name, obj):
        with open(os.path.join(out, name), "w") as f:
            json.dump(obj, f, indent=None, separators=(",", ":"))
            f.write("\n")

    dump("np.json", {
        "source": "data/NPannual.rda, data/NPpc.rda, data/theta.rda",
        "NPannual": {"year": np_annual["year"], "Qa": np_annual["Qa"]},
        "NPpc": {k: np_pc[k] for k in np_pc},
        "NPpc_first_year": 1200,
        "theta": theta_json(theta),
    })
    dump("p1.json", {
        "source": "R/sysdata.rda (P1annual, P1pc)",
        "P1annual": {"year": sysd["P1annual"]["year"], "Qa": sysd["P1annual"]["Qa"]},
        "P1pc": {k: sysd["P1pc"][k] for k in sysd["P1pc"]},
        "P1pc_first_year": 1600,
    })
    lds = sysd["NPlds"]
    dump("nplds.json", {
        "source": "R/sysdata.rda (NPlds = LDS_reconstruction(NPannual, t(NPpc), t(NPpc), start.year=1200))",
        "rec": {k: lds["rec"][k] for k in lds["rec"]},
        "theta": theta_json(lds["theta"]),
        "lik": mat1(lds["lik"])[0] if isinstance(mat1(lds["lik"]), list) else lds["lik"],
    })
    cv = sysd["NPcv"]
    md = cv["metrics.dist"]
    dump("npcv.json", {
        "source": "R/sysdata.rda (NPcv = cvLDS(...); unseeded random restarts: statistical anchor only)",
        "metrics_dist": {k: md[k] for k in md},
        "metrics": {k: cv["metrics"][k] for k in cv["metrics"]},
        "target": {k: cv["target"][k] for k in cv["target"]},
        "Ycv": {"year": cv["Ycv"]["year"], "Y": cv["Ycv"]["Y"]},
        "Z": cv["Z"],
    })
    # tests/testthat/test-LDS-EM.R:3-16 (inputs) and :26-40 (expected values, tolerance 1e-6 relative)
    dump("kat.json", {
        "source": "tests/testthat/test-LDS-EM.R:3-41",
        "theta0": {"A": 0.5, "B": [0.5] * 7, "C": 0.5, "D": [0.5] * 7, "Q": 1.0, "R": 1.0, "mu1": 1.0, "V1": 1.0},
        "P1pc_rows_1based": [322, 406],
        "tolerance": 1e-6,
        "smooth1_lik": -11.678657,
        "smooth1_X_1_85": [1.293356, -0.987671],
        "theta1": {"A": 0.606066, "C": -0.005995, "Q": 3.640236},
        "smooth2_lik": -0.114224,
        "theta2": {"A": 0.603945, "C": -0.012004, "Q": 3.644322},
        "em": {"niter": 100, "tol": 1e-5, "n_liks": 68, "lik": -0.039093},
    })
    print("wrote fixtures to", os.path.abspath(out))


if __name__ == "__main__":
    main()
