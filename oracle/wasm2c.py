#!/usr/bin/env python3
"""Mechanical WebAssembly(MVP) -> C translator used to build ``oracle/_ref``.

TEST INFRASTRUCTURE ONLY.  The reference's DSP engine exists solely as a compiled WASM blob embedded
as base64 in ``/root/reference/app/SignalsmithStretch.mjs:265`` (sha256 83869197...2d8ca3).  This tool
decodes that blob *where it lies* and emits one C function per wasm function (expression trees with
conservative spilling, ``goto`` for ``br``), so ``gcc -O2 -ffp-contract=off -fno-fast-math`` yields a
shared library that executes the reference bit-faithfully (wasm f32/f64 arithmetic is strict IEEE-754,
and the blob carries its own musl libm, so results do not depend on the host libm).

Nothing from the reference is stored in this repository: the generated C and the ``.so`` go to
``oracle/_ref/`` which is git-ignored (but travels to the GPU box with the snapshot).

Usage:  python oracle/wasm2c.py /root/reference/app/SignalsmithStretch.mjs oracle/_ref/stretch_ref.c
"""
import base64
import hashlib
import re
import struct
import sys

BLOB_SHA256 = "83869197b3c5ebf9fc8c517a1586aef1ecf77404842218d62b9c0e82882d8ca3"

I32, I64, F32, F64 = 0x7F, 0x7E, 0x7D, 0x7C
CT = {I32: "u32", I64: "u64", F32: "f32", F64: "f64"}


def extract_blob(mjs_path):
    """Pull the base64 wasm out of the .mjs (the data: URI on the line holding ``base64,``)."""
    text = open(mjs_path, "r", encoding="utf-8", errors="replace").read()
    best = None
    for m in re.finditer(r'base64,([A-Za-z0-9+/=]+)"', text):
        raw = base64.b64decode(m.group(1))
        if raw[:4] == b"\x00asm" and (best is None or len(raw) > len(best)):
            if hashlib.sha256(raw).hexdigest() == BLOB_SHA256:
                return raw
            best = raw
    if best is None:
        raise SystemExit("no wasm blob found in " + mjs_path)
    sys.stderr.write("warning: blob sha256 differs from the surveyed one\n")
    return best


class Reader:
    def __init__(self, data, pos=0, end=None):
        self.d = data
        self.p = pos
        self.end = len(data) if end is None else end

    def byte(self):
        b = self.d[self.p]
        self.p += 1
        return b

    def u(self):
        r = 0
        s = 0
        while True:
            b = self.byte()
            r |= (b & 0x7F) << s
            s += 7
            if not b & 0x80:
                return r

    def s(self, bits=64):
        r = 0
        s = 0
        while True:
            b = self.byte()
            r |= (b & 0x7F) << s
            s += 7
            if not b & 0x80:
                if b & 0x40:
                    r -= 1 << s
                return r

    def bytes(self, n):
        b = self.d[self.p:self.p + n]
        self.p += n
        return b

    def name(self):
        return self.bytes(self.u()).decode()

    def eof(self):
        return self.p >= self.end


class Module:
    pass


def parse(data):
    m = Module()
    m.types = []
    m.imports = []      # (mod, name, typeidx)
    m.funcs = []        # typeidx for defined functions
    m.table = []
    m.mem_min = 0
    m.globals = []      # (type, mut, init)
    m.exports = []      # (name, kind, idx)
    m.elems = []        # (offset, [funcidx])
    m.codes = []        # (locals, body_start, body_end)
    m.datas = []        # (offset, bytes)
    r = Reader(data)
    assert r.bytes(4) == b"\x00asm" and r.bytes(4) == b"\x01\x00\x00\x00"
    while not r.eof():
        sid = r.byte()
        size = r.u()
        end = r.p + size
        if sid == 1:
            for _ in range(r.u()):
                assert r.byte() == 0x60
                params = [r.byte() for _ in range(r.u())]
                results = [r.byte() for _ in range(r.u())]
                m.types.append((params, results))
        elif sid == 2:
            for _ in range(r.u()):
                mod, nm = r.name(), r.name()
                kind = r.byte()
                assert kind == 0, "only function imports supported"
                m.imports.append((mod, nm, r.u()))
        elif sid == 3:
            m.funcs = [r.u() for _ in range(r.u())]
        elif sid == 4:
            for _ in range(r.u()):
                assert r.byte() == 0x70
                flag = r.byte()
                mn = r.u()
                if flag & 1:
                    r.u()
                m.table_size = mn
        elif sid == 5:
            for _ in range(r.u()):
                flag = r.byte()
                m.mem_min = r.u()
                if flag & 1:
                    r.u()
        elif sid == 6:
            for _ in range(r.u()):
                t = r.byte()
                mut = r.byte()
                op = r.byte()
                assert op == 0x41
                v = r.s()
                assert r.byte() == 0x0B
                m.globals.append((t, mut, v))
        elif sid == 7:
            for _ in range(r.u()):
                nm = r.name()
                kind = r.byte()
                m.exports.append((nm, kind, r.u()))
        elif sid == 9:
            for _ in range(r.u()):
                flag = r.u()
                assert flag == 0
                assert r.byte() == 0x41
                off = r.s()
                assert r.byte() == 0x0B
                m.elems.append((off, [r.u() for _ in range(r.u())]))
        elif sid == 10:
            for _ in range(r.u()):
                bsize = r.u()
                bend = r.p + bsize
                locs = []
                for _ in range(r.u()):
                    n = r.u()
                    t = r.byte()
                    locs += [t] * n
                m.codes.append((locs, r.p, bend))
                r.p = bend
        elif sid == 11:
            for _ in range(r.u()):
                flag = r.u()
                assert flag == 0
                assert r.byte() == 0x41
                off = r.s()
                assert r.byte() == 0x0B
                m.datas.append((off, r.bytes(r.u())))
        r.p = end
    return m


# ---- expression model --------------------------------------------------------------------------
class E:
    __slots__ = ("s", "t", "locs", "mem", "glob")

    def __init__(self, s, t, locs=frozenset(), mem=False, glob=False):
        self.s = s
        self.t = t
        self.locs = locs
        self.mem = mem
        self.glob = glob


def comb(s, t, *es):
    locs = frozenset().union(*[e.locs for e in es]) if es else frozenset()
    return E(s, t, locs, any(e.mem for e in es), any(e.glob for e in es))


BIN_I = {  # name -> C template, a and b are unsigned of the op width
    "add": "({a} + {b})", "sub": "({a} - {b})", "mul": "({a} * {b})",
    "and": "({a} & {b})", "or": "({a} | {b})", "xor": "({a} ^ {b})",
}


class FuncGen:
    def __init__(self, mod, fidx, data, out):
        self.m = mod
        self.fidx = fidx
        self.d = data
        self.out = out
        self.ntmp = 0
        self.decls = []
        self.lines = []
        self.ind = 1

    # --- emit helpers
    def emit(self, s):
        self.lines.append("  " * self.ind + s)

    def tmp(self, t):
        n = "t%d" % self.ntmp
        self.ntmp += 1
        self.decls.append("%s %s;" % (CT[t], n))
        return n

    def spill_where(self, pred):
        for i, e in enumerate(self.st):
            if pred(e):
                n = self.tmp(e.t)
                self.emit("%s = %s;" % (n, e.s))
                self.st[i] = E(n, e.t)

    def spill_all(self):
        self.spill_where(lambda e: bool(e.locs) or e.mem or e.glob)

    def push(self, e):
        self.st.append(e)

    def pop(self):
        return self.st.pop()

    def functype(self, fi):
        ni = len(self.m.imports)
        ti = self.m.imports[fi][2] if fi < ni else self.m.funcs[fi - ni]
        return self.m.types[ti]

    def gen(self):
        m = self.m
        ni = len(m.imports)
        ti = m.funcs[self.fidx - ni]
        params, results = m.types[ti]
        locs, p0, p1 = m.codes[self.fidx - ni]
        self.ltypes = list(params) + list(locs)
        rett = CT[results[0]] if results else "void"
        sig = "static %s w%d(%s)" % (rett, self.fidx,
                                      ", ".join("%s l%d" % (CT[t], i) for i, t in enumerate(params)) or "void")
        r = Reader(self.d, p0, p1)
        self.st = []
        # control stack entries: dict(kind, label, res_t, res_var, height, is_loop)
        self.ctl = [dict(kind="func", label="Lret", res_t=(results[0] if results else None),
                         res_var=None, height=0, used=False)]
        if results:
            self.ctl[0]["res_var"] = "retv"
            self.decls.append("%s retv = 0;" % CT[results[0]])
        self.nlabel = 0
        self.run(r)
        o = self.out
        o.append(sig + " {")
        for i, t in enumerate(locs):
            o.append("  %s l%d = 0;" % (CT[t], i + len(params)))
        for dline in self.decls:
            o.append("  " + dline)
        o.extend(self.lines)
        o.append("  Lret:;")
        o.append("  return%s;" % (" retv" if results else ""))
        o.append("}")
        o.append("")

    def blocktype(self, r):
        b = r.byte()
        if b == 0x40:
            return None
        assert b in CT, "multi-value block types unsupported"
        return b

    def new_label(self):
        self.nlabel += 1
        return "L%d" % self.nlabel

    def skip_unreachable(self, r):
        """After an unconditional transfer: skip to the matching end/else of the current frame."""
        depth = 0
        while True:
            op = r.byte()
            if op in (0x02, 0x03, 0x04):
                r.byte() if self.d[r.p] in (0x40, 0x7F, 0x7E, 0x7D, 0x7C) else r.s()
                depth += 1
            elif op == 0x05:
                if depth == 0:
                    r.p -= 1
                    return
            elif op == 0x0B:
                if depth == 0:
                    r.p -= 1
                    return
                depth -= 1
            else:
                self.skip_imm(op, r)

    def skip_imm(self, op, r):
        if op in (0x0C, 0x0D, 0x10, 0x20, 0x21, 0x22, 0x23, 0x24):
            r.u()
        elif op == 0x0E:
            for _ in range(r.u() + 1):
                r.u()
        elif op == 0x11:
            r.u()
            r.u()
        elif 0x28 <= op <= 0x3E:
            r.u()
            r.u()
        elif op in (0x3F, 0x40):
            r.byte()
        elif op == 0x41:
            r.s()
        elif op == 0x42:
            r.s()
        elif op == 0x43:
            r.bytes(4)
        elif op == 0x44:
            r.bytes(8)

    def branch_to(self, depth):
        """Emit the transfer to ctl[-1-depth] (assign block result if any)."""
        fr = self.ctl[-1 - depth]
        fr["used"] = True
        s = ""
        if fr["kind"] != "loop" and fr["res_t"] is not None:
            s += "%s = %s; " % (fr["res_var"], self.st[-1].s)
        return s + "goto %s;" % fr["label"]

    def run(self, r):
        d = self.d
        while not r.eof():
            op = r.byte()
            # ---------------- control
            if op == 0x00:
                self.emit("wasm_trap(\"unreachable\");")
                self.dead(r)
            elif op == 0x01:
                pass
            elif op in (0x02, 0x03):
                bt = self.blocktype(r)
                self.spill_all()
                lab = self.new_label()
                fr = dict(kind="loop" if op == 0x03 else "block", label=lab, res_t=bt,
                          res_var=self.tmp(bt) if bt is not None else None, height=len(self.st), used=False)
                self.ctl.append(fr)
                if op == 0x03:
                    self.emit("%s:;" % lab)
                    self.emit("{")
                else:
                    self.emit("{")
                self.ind += 1
            elif op == 0x04:
                bt = self.blocktype(r)
                c = self.pop()
                self.spill_all()
                lab = self.new_label()
                fr = dict(kind="if", label=lab, res_t=bt, res_var=self.tmp(bt) if bt is not None else None,
                          height=len(self.st), used=False, has_else=False)
                self.ctl.append(fr)
                self.emit("if (%s) {" % c.s)
                self.ind += 1
            elif op == 0x05:
                fr = self.ctl[-1]
                if not fr.get("dead") and fr["res_t"] is not None:
                    self.emit("%s = %s;" % (fr["res_var"], self.pop().s))
                fr["dead"] = False
                del self.st[fr["height"]:]
                fr["has_else"] = True
                self.ind -= 1
                self.emit("} else {")
                self.ind += 1
            elif op == 0x0B:
                fr = self.ctl.pop()
                if fr["kind"] == "func":
                    if not fr.get("dead") and fr["res_t"] is not None:
                        self.emit("retv = %s;" % self.pop().s)
                    return
                if not fr.get("dead") and fr["res_t"] is not None:
                    self.emit("%s = %s;" % (fr["res_var"], self.pop().s))
                del self.st[fr["height"]:]
                self.ind -= 1
                self.emit("}")
                if fr["kind"] != "loop":
                    self.emit("%s:;" % fr["label"])
                if fr["res_t"] is not None:
                    self.push(E(fr["res_var"], fr["res_t"]))
            elif op == 0x0C:
                self.emit(self.branch_to(r.u()))
                self.dead(r)
            elif op == 0x0D:
                depth = r.u()
                c = self.pop()
                self.emit("if (%s) { %s }" % (c.s, self.branch_to(depth)))
            elif op == 0x0E:
                n = r.u()
                targets = [r.u() for _ in range(n + 1)]
                c = self.pop()
                self.emit("switch (%s) {" % c.s)
                for i, t in enumerate(targets[:-1]):
                    self.emit("  case %d: %s" % (i, self.branch_to(t)))
                self.emit("  default: %s" % self.branch_to(targets[-1]))
                self.emit("}")
                self.dead(r)
            elif op == 0x0F:
                self.emit(self.branch_to(len(self.ctl) - 1))
                self.dead(r)
            elif op == 0x10:
                fi = r.u()
                self.call("w%d" % fi if fi >= len(self.m.imports) else "imp_%s" % self.m.imports[fi][1],
                          self.functype(fi))
            elif op == 0x11:
                ti = r.u()
                r.u()
                idx = self.pop()
                self.call("call_indirect_%d" % ti, self.m.types[ti], extra=idx)
            elif op == 0x1A:
                self.pop()
            elif op == 0x1B:
                c = self.pop()
                b = self.pop()
                a = self.pop()
                self.push(comb("(%s ? %s : %s)" % (c.s, a.s, b.s), a.t, a, b, c))
            # ---------------- variables
            elif op == 0x20:
                k = r.u()
                self.push(E("l%d" % k, self.ltypes[k], frozenset([k])))
            elif op in (0x21, 0x22):
                k = r.u()
                v = self.pop()
                self.spill_where(lambda e: k in e.locs)
                self.emit("l%d = %s;" % (k, v.s))
                if op == 0x22:
                    self.push(E("l%d" % k, self.ltypes[k], frozenset([k])))
            elif op == 0x23:
                k = r.u()
                self.push(E("g%d" % k, self.m.globals[k][0], glob=True))
            elif op == 0x24:
                k = r.u()
                v = self.pop()
                self.spill_where(lambda e: e.glob)
                self.emit("g%d = %s;" % (k, v.s))
            # ---------------- memory
            elif 0x28 <= op <= 0x35:
                r.u()
                off = r.u()
                a = self.pop()
                fn, t = {0x28: ("ld_u32", I32), 0x29: ("ld_u64", I64), 0x2A: ("ld_f32", F32), 0x2B: ("ld_f64", F64),
                         0x2C: ("(u32)(int32_t)ld_s8", I32), 0x2D: ("(u32)ld_u8", I32),
                         0x2E: ("(u32)(int32_t)ld_s16", I32), 0x2F: ("(u32)ld_u16", I32),
                         0x30: ("(u64)(int64_t)ld_s8", I64), 0x31: ("(u64)ld_u8", I64),
                         0x32: ("(u64)(int64_t)ld_s16", I64), 0x33: ("(u64)ld_u16", I64),
                         0x34: ("(u64)(int64_t)(int32_t)ld_u32", I64), 0x35: ("(u64)ld_u32", I64)}[op]
                addr = a.s if off == 0 else "%s + %du" % (a.s, off)
                self.push(E("%s(%s)" % (fn, addr), t, a.locs, True, a.glob))
            elif 0x36 <= op <= 0x3E:
                r.u()
                off = r.u()
                v = self.pop()
                a = self.pop()
                fn = {0x36: "st_u32", 0x37: "st_u64", 0x38: "st_f32", 0x39: "st_f64", 0x3A: "st_u8", 0x3B: "st_u16",
                      0x3C: "st_u8", 0x3D: "st_u16", 0x3E: "st_u32"}[op]
                self.spill_where(lambda e: e.mem)
                addr = a.s if off == 0 else "%s + %du" % (a.s, off)
                self.emit("%s(%s, %s);" % (fn, addr, v.s))
            elif op == 0x3F:
                r.byte()
                self.push(E("wasm_pages", I32, glob=True))
            elif op == 0x40:
                r.byte()
                v = self.pop()
                self.spill_where(lambda e: e.mem or e.glob)
                n = self.tmp(I32)
                self.emit("%s = wasm_grow(%s);" % (n, v.s))
                self.push(E(n, I32))
            # ---------------- constants
            elif op == 0x41:
                self.push(E("%du" % (r.s() & 0xFFFFFFFF), I32))
            elif op == 0x42:
                self.push(E("%dull" % (r.s() & 0xFFFFFFFFFFFFFFFF), I64))
            elif op == 0x43:
                raw = r.bytes(4)
                self.push(E("f32c(0x%08xu)" % struct.unpack("<I", raw)[0], F32))
            elif op == 0x44:
                raw = r.bytes(8)
                self.push(E("f64c(0x%016xull)" % struct.unpack("<Q", raw)[0], F64))
            else:
                self.numeric(op)

    def dead(self, r):
        """Mark rest of the current frame unreachable and skip it."""
        self.ctl[-1]["dead"] = True
        self.skip_unreachable(r)

    def call(self, name, ftype, extra=None):
        params, results = ftype
        args = [self.pop() for _ in params][::-1]
        self.spill_where(lambda e: e.mem or e.glob)
        arglist = ", ".join(a.s for a in args)
        if extra is not None:
            arglist = extra.s + (", " + arglist if arglist else "")
        if results:
            n = self.tmp(results[0])
            self.emit("%s = %s(%s);" % (n, name, arglist))
            self.push(E(n, results[0]))
        else:
            self.emit("%s(%s);" % (name, arglist))

    def un(self, fmt, t):
        a = self.pop()
        self.push(comb(fmt.format(a=a.s), t, a))

    def bi(self, fmt, t):
        b = self.pop()
        a = self.pop()
        self.push(comb(fmt.format(a=a.s, b=b.s), t, a, b))

    def numeric(self, op):
        # i32 comparisons
        if op == 0x45:
            return self.un("(u32)({a} == 0)", I32)
        if op == 0x50:
            return self.un("(u32)({a} == 0)", I32)
        cmp_u = {0: "==", 1: "!=", 3: "<", 5: ">", 7: "<=", 9: ">="}
        cmp_s = {2: "<", 4: ">", 6: "<=", 8: ">="}
        if 0x46 <= op <= 0x4F:
            k = op - 0x46
            if k in cmp_u:
                return self.bi("(u32)({a} %s {b})" % cmp_u[k], I32)
            return self.bi("(u32)((int32_t){a} %s (int32_t){b})" % cmp_s[k], I32)
        if 0x51 <= op <= 0x5A:
            k = op - 0x51
            if k in cmp_u:
                return self.bi("(u32)({a} %s {b})" % cmp_u[k], I32)
            return self.bi("(u32)((int64_t){a} %s (int64_t){b})" % cmp_s[k], I32)
        fcmp = ["==", "!=", "<", ">", "<=", ">="]
        if 0x5B <= op <= 0x60:
            return self.bi("(u32)({a} %s {b})" % fcmp[op - 0x5B], I32)
        if 0x61 <= op <= 0x66:
            return self.bi("(u32)({a} %s {b})" % fcmp[op - 0x61], I32)
        # integer arithmetic
        for base, t, w, st in ((0x67, I32, 32, "int32_t"), (0x79, I64, 64, "int64_t")):
            if base <= op <= base + 17:
                k = op - base
                ut = CT[t]
                if k == 0:
                    return self.un("(%s)wasm_clz%d({a})" % (ut, w), t)
                if k == 1:
                    return self.un("(%s)wasm_ctz%d({a})" % (ut, w), t)
                if k == 2:
                    return self.un("(%s)__builtin_popcount%s({a})" % (ut, "ll" if w == 64 else ""), t)
                names = ["add", "sub", "mul", "div_s", "div_u", "rem_s", "rem_u", "and", "or", "xor", "shl", "shr_s",
                         "shr_u", "rotl", "rotr"]
                nm = names[k - 3]
                if nm in BIN_I:
                    return self.bi("(%s)%s" % (ut, BIN_I[nm]), t)
                if nm == "div_s":
                    return self.bi("(%s)((%s){a} / (%s){b})" % (ut, st, st), t)
                if nm == "div_u":
                    return self.bi("(%s)({a} / {b})" % ut, t)
                if nm == "rem_s":
                    return self.bi("(%s)wasm_rem_s%d({a}, {b})" % (ut, w), t)
                if nm == "rem_u":
                    return self.bi("(%s)({a} %% {b})" % ut, t)
                if nm == "shl":
                    return self.bi("(%s)({a} << ({b} & %d))" % (ut, w - 1), t)
                if nm == "shr_s":
                    return self.bi("(%s)((%s){a} >> ({b} & %d))" % (ut, st, w - 1), t)
                if nm == "shr_u":
                    return self.bi("(%s)({a} >> ({b} & %d))" % (ut, w - 1), t)
                if nm == "rotl":
                    return self.bi("(%s)wasm_rotl%d({a}, {b})" % (ut, w), t)
                if nm == "rotr":
                    return self.bi("(%s)wasm_rotr%d({a}, {b})" % (ut, w), t)
        # float arithmetic
        for base, t, sfx in ((0x8B, F32, "f"), (0x99, F64, "")):
            if base <= op <= base + 13:
                k = op - base
                un = {0: "fabs%s({a})", 1: "(-{a})", 2: "ceil%s({a})", 3: "floor%s({a})", 4: "trunc%s({a})",
                      5: "nearbyint%s({a})", 6: "sqrt%s({a})"}
                if k in un:
                    f = un[k]
                    return self.un(f % sfx if "%s" in f else f, t)
                bi = {7: "({a} + {b})", 8: "({a} - {b})", 9: "({a} * {b})", 10: "({a} / {b})",
                      11: "wasm_fmin%s({a}, {b})", 12: "wasm_fmax%s({a}, {b})", 13: "copysign%s({a}, {b})"}
                f = bi[k]
                return self.bi(f % sfx if "%s" in f else f, t)
        conv = {
            0xA7: ("(u32){a}", I32),                       # i32.wrap_i64
            0xA8: ("(u32)(int32_t){a}", I32), 0xA9: ("(u32){a}", I32),     # trunc f32 s/u
            0xAA: ("(u32)(int32_t){a}", I32), 0xAB: ("(u32){a}", I32),     # trunc f64 s/u
            0xAC: ("(u64)(int64_t)(int32_t){a}", I64), 0xAD: ("(u64){a}", I64),
            0xAE: ("(u64)(int64_t){a}", I64), 0xAF: ("(u64){a}", I64),
            0xB0: ("(u64)(int64_t){a}", I64), 0xB1: ("(u64){a}", I64),
            0xB2: ("(f32)(int32_t){a}", F32), 0xB3: ("(f32){a}", F32),
            0xB4: ("(f32)(int64_t){a}", F32), 0xB5: ("(f32){a}", F32),
            0xB6: ("(f32){a}", F32),
            0xB7: ("(f64)(int32_t){a}", F64), 0xB8: ("(f64){a}", F64),
            0xB9: ("(f64)(int64_t){a}", F64), 0xBA: ("(f64){a}", F64),
            0xBB: ("(f64){a}", F64),
            0xBC: ("f32_bits({a})", I32), 0xBD: ("f64_bits({a})", I64),
            0xBE: ("bits_f32({a})", F32), 0xBF: ("bits_f64({a})", F64),
            0xC0: ("(u32)(int32_t)(int8_t){a}", I32), 0xC1: ("(u32)(int32_t)(int16_t){a}", I32),
            0xC2: ("(u64)(int64_t)(int8_t){a}", I64), 0xC3: ("(u64)(int64_t)(int16_t){a}", I64),
            0xC4: ("(u64)(int64_t)(int32_t){a}", I64),
        }
        if op in conv:
            f, t = conv[op]
            return self.un(f, t)
        raise SystemExit("unsupported opcode 0x%02x in f%d" % (op, self.fidx))


PRELUDE = r'''/* GENERATED by oracle/wasm2c.py from the reference's embedded wasm blob -- do not commit. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <stdio.h>
typedef uint32_t u32; typedef uint64_t u64; typedef float f32; typedef double f64;

typedef struct wasm_inst { uint8_t *mem; u32 pages; u32 max_pages; u32 g[8]; u32 rng_seed; } wasm_inst;
static __thread wasm_inst *cur;
#define wasm_pages (cur->pages)
#define MEM (cur->mem)
static void wasm_trap(const char *why) { fprintf(stderr, "wasm trap: %s\n", why); abort(); }
static inline f32 f32c(u32 b) { f32 f; memcpy(&f, &b, 4); return f; }
static inline f64 f64c(u64 b) { f64 f; memcpy(&f, &b, 8); return f; }
static inline u32 f32_bits(f32 f) { u32 b; memcpy(&b, &f, 4); return b; }
static inline u64 f64_bits(f64 f) { u64 b; memcpy(&b, &f, 8); return b; }
#define bits_f32 f32c
#define bits_f64 f64c
#define LD(T, name) static inline T name(u32 a) { T v; memcpy(&v, MEM + a, sizeof(T)); return v; }
LD(u32, ld_u32) LD(u64, ld_u64) LD(f32, ld_f32) LD(f64, ld_f64) LD(uint8_t, ld_u8) LD(int8_t, ld_s8)
LD(uint16_t, ld_u16) LD(int16_t, ld_s16)
#define ST(T, name) static inline void name(u32 a, T v) { memcpy(MEM + a, &v, sizeof(T)); }
ST(u32, st_u32) ST(u64, st_u64) ST(f32, st_f32) ST(f64, st_f64) ST(uint8_t, st_u8) ST(uint16_t, st_u16)
static inline u32 wasm_clz32(u32 x) { return x ? __builtin_clz(x) : 32; }
static inline u32 wasm_ctz32(u32 x) { return x ? __builtin_ctz(x) : 32; }
static inline u64 wasm_clz64(u64 x) { return x ? __builtin_clzll(x) : 64; }
static inline u64 wasm_ctz64(u64 x) { return x ? __builtin_ctzll(x) : 64; }
static inline u32 wasm_rotl32(u32 x, u32 n) { n &= 31; return n ? (x << n) | (x >> (32 - n)) : x; }
static inline u32 wasm_rotr32(u32 x, u32 n) { n &= 31; return n ? (x >> n) | (x << (32 - n)) : x; }
static inline u64 wasm_rotl64(u64 x, u64 n) { n &= 63; return n ? (x << n) | (x >> (64 - n)) : x; }
static inline u64 wasm_rotr64(u64 x, u64 n) { n &= 63; return n ? (x >> n) | (x << (64 - n)) : x; }
static inline int32_t wasm_rem_s32(u32 a, u32 b) { return (int32_t)b == -1 ? 0 : (int32_t)a % (int32_t)b; }
static inline int64_t wasm_rem_s64(u64 a, u64 b) { return (int64_t)b == -1 ? 0 : (int64_t)a % (int64_t)b; }
static inline f32 wasm_fminf(f32 a, f32 b) { return (a != a || b != b) ? NAN : fminf(a, b); }
static inline f32 wasm_fmaxf(f32 a, f32 b) { return (a != a || b != b) ? NAN : fmaxf(a, b); }
static inline f64 wasm_fmin(f64 a, f64 b) { return (a != a || b != b) ? NAN : fmin(a, b); }
static inline f64 wasm_fmax(f64 a, f64 b) { return (a != a || b != b) ? NAN : fmax(a, b); }
static u32 wasm_grow(u32 n) { u32 old = cur->pages; if (old + n > cur->max_pages) return (u32)-1; cur->pages = old + n; return old; }
'''


def generate(data):
    m = parse(data)
    ni = len(m.imports)
    out = [PRELUDE]
    for gi, (t, mut, v) in enumerate(m.globals):
        out.append("#define g%d (cur->g[%d])" % (gi, gi))
    # imports (emscripten contract, app/SignalsmithStretch.mjs:454-459)
    out.append(r'''
/* a.a random_get(buf,len): seedable instead of crypto.getRandomValues */
static u32 imp_a(u32 buf, u32 len) { u32 s = cur->rng_seed; for (u32 i = 0; i < len; ++i) MEM[buf + i] = (uint8_t)(s >> (8 * (i & 3))); return 0; }
/* a.b emscripten_resize_heap(requestedSize) */
static u32 imp_b(u32 req) { u32 need = (req + 65535u) >> 16; if (need > cur->max_pages) return 0; if (need > cur->pages) cur->pages = need; return 1; }
/* a.c emscripten_memcpy_js(dest, src, num) */
static void imp_c(u32 d, u32 s, u32 n) { memmove(MEM + d, MEM + s, n); }
/* a.d abort */
static void imp_d(void) { wasm_trap("abort()"); }
''')
    # prototypes
    for i, ti in enumerate(m.funcs):
        params, results = m.types[ti]
        out.append("static %s w%d(%s);" % (CT[results[0]] if results else "void", i + ni,
                                           ", ".join(CT[t] for t in params) or "void"))
    # call_indirect dispatchers per type
    table = {}
    for off, fl in m.elems:
        for k, fi in enumerate(fl):
            table[off + k] = fi
    for ti, (params, results) in enumerate(m.types):
        rett = CT[results[0]] if results else "void"
        args = ", ".join("%s a%d" % (CT[t], i) for i, t in enumerate(params))
        out.append("static %s call_indirect_%d(u32 idx%s) {" % (rett, ti, (", " + args) if args else ""))
        out.append("  switch (idx) {")
        for slot, fi in sorted(table.items()):
            fti = m.imports[fi][2] if fi < ni else m.funcs[fi - ni]
            if m.types[fti] == (params, results):
                nm = "w%d" % fi if fi >= ni else "imp_%s" % m.imports[fi][1]
                call = "%s(%s)" % (nm, ", ".join("a%d" % i for i in range(len(params))))
                out.append("    case %d: %s" % (slot, ("return " + call + ";") if results else (call + "; return;")))
        out.append("  }")
        out.append("  wasm_trap(\"call_indirect\");%s" % (" return 0;" if results else ""))
        out.append("}")
    out.append("")
    for i in range(len(m.funcs)):
        FuncGen(m, i + ni, data, out).gen()
    # instance management + exports
    for k, (off, b) in enumerate(m.datas):
        out.append("static const uint8_t wasm_data_%d[] = {%s};" % (k, ",".join(str(x) for x in b)))
    out.append("static const struct { u32 off; u32 len; const uint8_t *bytes; } wasm_data[] = {")
    for k, (off, b) in enumerate(m.datas):
        out.append("  {%du, %du, wasm_data_%d}," % (off, len(b), k))
    out.append("};")
    exp = {nm: idx for nm, kind, idx in m.exports if kind == 0}
    ginit = "".join("  w->g[%d] = %du;\n" % (i, v & 0xFFFFFFFF) for i, (t, mut, v) in enumerate(m.globals))
    out.append(r'''
#define API __attribute__((visibility("default")))
API wasm_inst *ref_new(u32 seed) {
  wasm_inst *w = (wasm_inst *)calloc(1, sizeof(wasm_inst));
  w->max_pages = 4096; /* 256 MiB reserve, lazily committed */
  w->mem = (uint8_t *)calloc((size_t)w->max_pages, 65536);
  w->pages = %du;
  w->rng_seed = seed;
%s  for (unsigned k = 0; k < sizeof(wasm_data) / sizeof(wasm_data[0]); ++k) memcpy(w->mem + wasm_data[k].off, wasm_data[k].bytes, wasm_data[k].len);
  cur = w;
  w%d(); /* __wasm_call_ctors */
  w%d(0, 0); /* main */
  return w;
}
API void ref_free(wasm_inst *w) { if (cur == w) cur = 0; free(w->mem); free(w); }
API void ref_select(wasm_inst *w) { cur = w; }
API uint8_t *ref_memory(wasm_inst *w) { return w->mem; }
API u32 ref_memory_bytes(wasm_inst *w) { return w->pages << 16; }
''' % (m.mem_min, ginit, exp["f"], exp["y"]))
    # export letter -> name table, app/SignalsmithStretch.mjs:462-479
    names = dict(h="setBuffers", i="blockSamples", j="intervalSamples", k="inputLatency", l="outputLatency",
                 m="reset", n="presetDefault", o="presetCheaper", p="configure", q="setTransposeFactor",
                 r="setTransposeSemitones", s="setFormantFactor", t="setFormantSemitones", u="setFormantBase",
                 v="seek", w="process", x="flush")
    for letter, nm in names.items():
        fi = exp[letter]
        params, results = m.types[m.funcs[fi - ni]]
        rett = CT[results[0]] if results else "void"
        args = ", ".join("%s a%d" % (CT[t], i) for i, t in enumerate(params))
        call = "w%d(%s)" % (fi, ", ".join("a%d" % i for i in range(len(params))))
        out.append("API %s ref_%s(%s) { %s%s; }" % (rett, nm, args or "void", "return " if results else "", call))
    # map of wasm function index -> export name, for humans reading the generated C
    out.append("/* exports: " + ", ".join("%s=w%d" % (names.get(n, n), i) for n, k, i in m.exports if k == 0) + " */")
    return "\n".join(out) + "\n"


if __name__ == "__main__":
    src, dst = sys.argv[1], sys.argv[2]
    code = generate(extract_blob(src))
    open(dst, "w").write(code)
