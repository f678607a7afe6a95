#!/usr/bin/env python3
"""wasm2c.py -- ahead-of-time translator: WebAssembly (MVP + sign-ext, sat-trunc,
bulk-memory copy/fill, multi-value) -> portable C.

TEST INFRASTRUCTURE ONLY (oracle/).  Purpose: the reference ships its stwo prover
only as Rust source (no Rust toolchain here) and as the compiled artefact
`/root/reference/resources/stwo/s2circuits_bg.wasm`.  This script
turns that artefact into C under `oracle/_ref/` (git-ignored, never committed), which
`oracle/Makefile` compiles into `oracle/_ref/libs2c_ref.so`.  That shared object *is*
the reference implementation (pinned stwo rev f117d487, Blake2sMerkleChannel,
PcsConfig::default) running natively, and is what pins the CPU restatement in
`oracle/` and the CUDA product to reference proof bytes.

Nothing here is derived from reference source text: it is a generic wasm compiler.

Usage: wasm2c.py in.wasm outdir [n_chunks]
"""
import struct
import sys
import os

I32, I64, F32, F64 = 0x7F, 0x7E, 0x7D, 0x7C
CT = {I32: "u32", I64: "u64", F32: "f32", F64: "f64"}
SUF = {I32: "i", I64: "j", F32: "f", F64: "d"}


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
            b = self.d[self.p]
            self.p += 1
            r |= (b & 0x7F) << s
            s += 7
            if not b & 0x80:
                return r

    def s(self, bits):
        r = 0
        s = 0
        while True:
            b = self.d[self.p]
            self.p += 1
            r |= (b & 0x7F) << s
            s += 7
            if not b & 0x80:
                if b & 0x40 and s < bits + 7:
                    r |= -1 << s
                return r

    def bytes(self, n):
        b = self.d[self.p:self.p + n]
        self.p += n
        return b

    def name(self):
        return self.bytes(self.u()).decode()


class Module:
    pass


def parse(data):
    m = Module()
    m.types = []
    m.imports = []      # (mod, name, kind, desc)
    m.func_types = []   # type index per defined function
    m.tables = []
    m.mems = []
    m.globals = []      # (type, mut, init_expr)
    m.exports = []
    m.elems = []
    m.codes = []
    m.datas = []
    m.start = None
    r = Reader(data, 8)
    while r.p < len(data):
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
                mod = r.name()
                nm = r.name()
                kind = r.byte()
                if kind == 0:
                    desc = r.u()
                elif kind == 1:
                    rt = r.byte()
                    fl = r.byte()
                    desc = (rt, r.u(), r.u() if fl & 1 else None)
                elif kind == 2:
                    fl = r.byte()
                    desc = (r.u(), r.u() if fl & 1 else None)
                else:
                    desc = (r.byte(), r.byte())
                m.imports.append((mod, nm, kind, desc))
        elif sid == 3:
            m.func_types = [r.u() for _ in range(r.u())]
        elif sid == 4:
            for _ in range(r.u()):
                rt = r.byte()
                fl = r.byte()
                m.tables.append((rt, r.u(), r.u() if fl & 1 else None))
        elif sid == 5:
            for _ in range(r.u()):
                fl = r.byte()
                m.mems.append((r.u(), r.u() if fl & 1 else None))
        elif sid == 6:
            for _ in range(r.u()):
                t = r.byte()
                mut = r.byte()
                m.globals.append((t, mut, const_expr(r)))
        elif sid == 7:
            for _ in range(r.u()):
                nm = r.name()
                kind = r.byte()
                m.exports.append((nm, kind, r.u()))
        elif sid == 8:
            m.start = r.u()
        elif sid == 9:
            for _ in range(r.u()):
                flag = r.u()
                if flag == 0:
                    off = const_expr(r)
                    m.elems.append((0, off, [r.u() for _ in range(r.u())]))
                elif flag == 2:
                    tbl = r.u()
                    off = const_expr(r)
                    assert r.byte() == 0
                    m.elems.append((tbl, off, [r.u() for _ in range(r.u())]))
                else:
                    raise NotImplementedError("elem flag %d" % flag)
        elif sid == 10:
            for _ in range(r.u()):
                sz = r.u()
                fend = r.p + sz
                locs = []
                for _ in range(r.u()):
                    n = r.u()
                    t = r.byte()
                    locs.append((n, t))
                m.codes.append((locs, r.p, fend))
                r.p = fend
        elif sid == 11:
            for _ in range(r.u()):
                flag = r.u()
                if flag == 0:
                    off = const_expr(r)
                    m.datas.append((off, r.bytes(r.u())))
                elif flag == 1:
                    m.datas.append((None, r.bytes(r.u())))
                else:
                    raise NotImplementedError
        r.p = end
    return m


def const_expr(r):
    op = r.byte()
    if op == 0x41:
        v = ("i32", r.s(32))
    elif op == 0x42:
        v = ("i64", r.s(64))
    elif op == 0x23:
        v = ("global", r.u())
    elif op == 0xD0:
        r.byte()
        v = ("i32", 0)
    else:
        raise NotImplementedError("const op %x" % op)
    assert r.byte() == 0x0B
    return v


BIN_I = {  # opcode offset from 0x6A (i32) / 0x7C (i64)
    0: "+", 1: "-", 2: "*", 7: "&", 8: "|", 9: "^",
}
CMP_U = {0x46: "==", 0x47: "!=", 0x49: "<", 0x4B: ">", 0x4D: "<=", 0x4F: ">="}
CMP_S = {0x48: "<", 0x4A: ">", 0x4C: "<=", 0x4E: ">="}
CMP64_U = {0x51: "==", 0x52: "!=", 0x54: "<", 0x56: ">", 0x58: "<=", 0x5A: ">="}
CMP64_S = {0x53: "<", 0x55: ">", 0x57: "<=", 0x59: ">="}
FCMP = {0: "==", 1: "!=", 2: "<", 3: ">", 4: "<=", 5: ">="}

LOADS = {
    0x28: (I32, "u32", None), 0x29: (I64, "u64", None), 0x2A: (F32, "f32", None), 0x2B: (F64, "f64", None),
    0x2C: (I32, "int8_t", "(u32)(int32_t)"), 0x2D: (I32, "uint8_t", "(u32)"),
    0x2E: (I32, "int16_t", "(u32)(int32_t)"), 0x2F: (I32, "uint16_t", "(u32)"),
    0x30: (I64, "int8_t", "(u64)(int64_t)"), 0x31: (I64, "uint8_t", "(u64)"),
    0x32: (I64, "int16_t", "(u64)(int64_t)"), 0x33: (I64, "uint16_t", "(u64)"),
    0x34: (I64, "int32_t", "(u64)(int64_t)"), 0x35: (I64, "uint32_t", "(u64)"),
}
STORES = {
    0x36: (I32, "u32"), 0x37: (I64, "u64"), 0x38: (F32, "f32"), 0x39: (F64, "f64"),
    0x3A: (I32, "uint8_t"), 0x3B: (I32, "uint16_t"),
    0x3C: (I64, "uint8_t"), 0x3D: (I64, "uint16_t"), 0x3E: (I64, "uint32_t"),
}


class Frame:
    __slots__ = ("kind", "label", "height", "params", "results", "dead", "used", "entry_dead")


class FuncGen:
    def __init__(self, mod, fidx, nimp):
        self.m = mod
        self.fidx = fidx
        self.nimp = nimp
        self.out = []
        self.stack = []      # list of types
        self.maxdepth = {}   # (depth,type) used
        self.nlabel = 0

    def sv(self, depth, t):
        self.maxdepth[(depth, t)] = True
        return "s%d%s" % (depth, SUF[t])

    def push(self, t):
        self.stack.append(t)
        return self.sv(len(self.stack) - 1, t)

    def pop(self, t=None):
        tt = self.stack.pop()
        if t is not None:
            assert tt == t, "type mismatch in f%d: want %x got %x" % (self.fidx, t, tt)
        return self.sv(len(self.stack), tt)

    def top(self):
        return self.sv(len(self.stack) - 1, self.stack[-1])

    def emit(self, s):
        self.out.append(s)

    def blocktype(self, r):
        b = self.m.data[r.p]
        if b == 0x40:
            r.p += 1
            return [], []
        if b in CT:
            r.p += 1
            return [], [b]
        ti = r.s(33)
        p, q = self.m.types[ti]
        return list(p), list(q)

    def func_type(self, f):
        if f < self.nimp:
            return self.m.types[self.m.imports_f[f]]
        return self.m.types[self.m.func_types[f - self.nimp]]

    def branch_moves(self, fr):
        """Emit moves of the top-N values to the target frame's base; returns code."""
        tys = fr.params if fr.kind == "loop" else fr.results
        n = len(tys)
        src0 = len(self.stack) - n
        code = []
        if src0 != fr.height:
            for i, t in enumerate(tys):
                assert self.stack[src0 + i] == t
                code.append("%s=%s;" % (self.sv(fr.height + i, t), self.sv(src0 + i, t)))
        fr.used = True
        return "".join(code)

    def callsig(self, ty):
        return ty

    def gen(self):
        m = self.m
        locs, start, end = m.codes[self.fidx - self.nimp]
        params, results = self.func_type(self.fidx)
        ltypes = list(params)
        for n, t in locs:
            ltypes += [t] * n
        self.ltypes = ltypes
        r = Reader(m.data, start, end)
        ctrl = []
        fr = Frame()
        fr.kind = "func"
        fr.label = self.nlabel
        self.nlabel += 1
        fr.height = 0
        fr.params = []
        fr.results = list(results)
        fr.dead = False
        fr.used = False
        fr.entry_dead = False
        ctrl.append(fr)
        E = self.emit
        while r.p < end:
            op = r.byte()
            cur = ctrl[-1]
            dead = cur.dead
            # ---- control ----
            if op == 0x00:
                if not dead:
                    E("TRAP(%d);" % self.fidx)
                    cur.dead = True
            elif op == 0x01:
                pass
            elif op in (0x02, 0x03, 0x04):
                p, q = self.blocktype(r)
                f = Frame()
                f.kind = {2: "block", 3: "loop", 4: "if"}[op]
                f.label = self.nlabel
                self.nlabel += 1
                f.params = p
                f.results = q
                f.dead = dead
                f.entry_dead = dead
                f.used = False
                if not dead:
                    if op == 4:
                        c = self.pop(I32)
                    f.height = len(self.stack) - len(p)
                    if op == 3:
                        E("L%d:;" % f.label)
                    if op == 4:
                        E("if(%s){" % c)
                else:
                    f.height = 0
                ctrl.append(f)
            elif op == 0x05:  # else
                f = cur
                if not f.entry_dead:
                    if not f.dead:
                        assert len(self.stack) == f.height + len(f.results)
                    E("}else{")
                    self.stack = self.stack[:f.height] + list(f.params)
                    f.dead = False
                    f.kind = "else"
            elif op == 0x0B:  # end
                f = ctrl.pop()
                if not f.entry_dead:
                    if not f.dead:
                        assert len(self.stack) == f.height + len(f.results), \
                            "stack height f%d: %d vs %d+%d" % (self.fidx, len(self.stack), f.height, len(f.results))
                    if f.kind in ("if", "else"):
                        if f.kind == "if":
                            assert f.params == f.results
                        E("}")
                    if f.kind == "func":
                        if not f.dead or f.used:
                            if f.used:
                                E("L%d:;" % f.label)
                            self.stack = list(f.results)
                            self.emit_return(results)
                    else:
                        if f.kind != "loop" and f.used:
                            E("L%d:;" % f.label)
                        self.stack = self.stack[:f.height] + list(f.results)
                        # code after the block is reachable only if fallthrough or branched to
                        if f.dead and not f.used and f.kind not in ("if",):
                            ctrl[-1].dead = True
                        if f.kind == "else" and f.dead and not f.used:
                            # both arms may be dead only if 'then' arm was also dead; be conservative
                            ctrl[-1].dead = False
            elif op == 0x0C:
                d = r.u()
                if not dead:
                    t = ctrl[-1 - d]
                    if t.kind == "func":
                        self.emit_return(results)
                    else:
                        E(self.branch_moves(t) + "goto L%d;" % t.label)
                    cur.dead = True
            elif op == 0x0D:
                d = r.u()
                if not dead:
                    c = self.pop(I32)
                    t = ctrl[-1 - d]
                    if t.kind == "func":
                        E("if(%s){" % c)
                        self.emit_return(results)
                        E("}")
                    else:
                        E("if(%s){%sgoto L%d;}" % (c, self.branch_moves(t), t.label))
            elif op == 0x0E:
                n = r.u()
                tg = [r.u() for _ in range(n + 1)]
                if not dead:
                    c = self.pop(I32)
                    E("switch(%s){" % c)
                    for i, d in enumerate(tg):
                        t = ctrl[-1 - d]
                        lab = "default" if i == n else "case %d" % i
                        if t.kind == "func":
                            E("%s:" % lab)
                            self.emit_return(results)
                        else:
                            E("%s:%sgoto L%d;" % (lab, self.branch_moves(t), t.label))
                    E("}")
                    cur.dead = True
            elif op == 0x0F:
                if not dead:
                    self.emit_return(results)
                    cur.dead = True
            elif op == 0x10:
                f = r.u()
                if not dead:
                    self.emit_call("fn%d" % f, self.func_type(f))
            elif op == 0x11:
                ti = r.u()
                tbl = r.u()
                if not dead:
                    idx = self.pop(I32)
                    p, q = m.types[ti]
                    cast = "((%s(*)(%s))tbl_get(%s))" % (
                        rettype(q), ",".join(CT[t] for t in p) or "void", idx)
                    self.emit_call(cast, (p, q))
            elif op == 0x1A:
                if not dead:
                    self.pop()
            elif op in (0x1B, 0x1C):
                if op == 0x1C:
                    for _ in range(r.u()):
                        r.byte()
                if not dead:
                    c = self.pop(I32)
                    b = self.pop()
                    t = self.stack[-1]
                    a = self.top()
                    E("%s=%s?%s:%s;" % (a, c, a, b))
            # ---- variables ----
            elif op == 0x20:
                i = r.u()
                if not dead:
                    E("%s=l%d;" % (self.push(ltypes[i]), i))
            elif op == 0x21:
                i = r.u()
                if not dead:
                    E("l%d=%s;" % (i, self.pop(ltypes[i])))
            elif op == 0x22:
                i = r.u()
                if not dead:
                    assert self.stack[-1] == ltypes[i]
                    E("l%d=%s;" % (i, self.top()))
            elif op == 0x23:
                i = r.u()
                if not dead:
                    E("%s=g%d;" % (self.push(m.globals[i][0]), i))
            elif op == 0x24:
                i = r.u()
                if not dead:
                    E("g%d=%s;" % (i, self.pop(m.globals[i][0])))
            # ---- memory ----
            elif op in LOADS:
                r.u()
                off = r.u()
                if not dead:
                    t, ct, cast = LOADS[op]
                    a = self.pop(I32)
                    d = self.push(t)
                    E("{%s t_;memcpy(&t_,mem+(u64)%s+%du,sizeof t_);%s=%st_;}" % (ct, a, off, d, cast or ""))
            elif op in STORES:
                r.u()
                off = r.u()
                if not dead:
                    t, ct = STORES[op]
                    v = self.pop(t)
                    a = self.pop(I32)
                    E("{%s t_=(%s)%s;memcpy(mem+(u64)%s+%du,&t_,sizeof t_);}" % (ct, ct, v, a, off))
            elif op == 0x3F:
                r.byte()
                if not dead:
                    E("%s=mem_pages;" % self.push(I32))
            elif op == 0x40:
                r.byte()
                if not dead:
                    a = self.top()
                    E("%s=mem_grow(%s);" % (a, a))
            # ---- consts ----
            elif op == 0x41:
                v = r.s(32)
                if not dead:
                    E("%s=%du;" % (self.push(I32), v & 0xFFFFFFFF))
            elif op == 0x42:
                v = r.s(64)
                if not dead:
                    E("%s=%dull;" % (self.push(I64), v & 0xFFFFFFFFFFFFFFFF))
            elif op == 0x43:
                b = r.bytes(4)
                if not dead:
                    E("%s=f32_bits(%du);" % (self.push(F32), struct.unpack("<I", b)[0]))
            elif op == 0x44:
                b = r.bytes(8)
                if not dead:
                    E("%s=f64_bits(%dull);" % (self.push(F64), struct.unpack("<Q", b)[0]))
            elif dead:
                self.skip_imm(op, r)
            # ---- i32 compare ----
            elif op == 0x45:
                a = self.top()
                assert self.stack[-1] == I32
                E("%s=!%s;" % (a, a))
            elif op in CMP_U:
                b = self.pop(I32); a = self.top()
                E("%s=%s%s%s;" % (a, a, CMP_U[op], b))
            elif op in CMP_S:
                b = self.pop(I32); a = self.top()
                E("%s=(int32_t)%s%s(int32_t)%s;" % (a, a, CMP_S[op], b))
            elif op == 0x50:
                a = self.pop(I64)
                E("%s=!%s;" % (self.push(I32), a))
            elif op in CMP64_U:
                b = self.pop(I64); a = self.pop(I64)
                E("%s=%s%s%s;" % (self.push(I32), a, CMP64_U[op], b))
            elif op in CMP64_S:
                b = self.pop(I64); a = self.pop(I64)
                E("%s=(int64_t)%s%s(int64_t)%s;" % (self.push(I32), a, CMP64_S[op], b))
            elif 0x5B <= op <= 0x60:
                b = self.pop(F32); a = self.pop(F32)
                E("%s=%s%s%s;" % (self.push(I32), a, FCMP[op - 0x5B], b))
            elif 0x61 <= op <= 0x66:
                b = self.pop(F64); a = self.pop(F64)
                E("%s=%s%s%s;" % (self.push(I32), a, FCMP[op - 0x61], b))
            # ---- i32 arith ----
            elif 0x67 <= op <= 0x69:
                a = self.top()
                fn = ["clz32", "ctz32", "popcnt32"][op - 0x67]
                E("%s=%s(%s);" % (a, fn, a))
            elif 0x6A <= op <= 0x78:
                b = self.pop(I32); a = self.top()
                k = op - 0x6A
                if k in BIN_I:
                    E("%s=%s%s%s;" % (a, a, BIN_I[k], b))
                elif k == 3:
                    E("%s=div_s32(%s,%s);" % (a, a, b))
                elif k == 4:
                    E("if(!%s)TRAP(%d);%s=%s/%s;" % (b, self.fidx, a, a, b))
                elif k == 5:
                    E("%s=rem_s32(%s,%s);" % (a, a, b))
                elif k == 6:
                    E("if(!%s)TRAP(%d);%s=%s%%%s;" % (b, self.fidx, a, a, b))
                elif k == 10:
                    E("%s=%s<<(%s&31);" % (a, a, b))
                elif k == 11:
                    E("%s=(u32)((int32_t)%s>>(%s&31));" % (a, a, b))
                elif k == 12:
                    E("%s=%s>>(%s&31);" % (a, a, b))
                elif k == 13:
                    E("%s=rotl32(%s,%s);" % (a, a, b))
                elif k == 14:
                    E("%s=rotr32(%s,%s);" % (a, a, b))
            elif 0x79 <= op <= 0x7B:
                a = self.top()
                fn = ["clz64", "ctz64", "popcnt64"][op - 0x79]
                E("%s=%s(%s);" % (a, fn, a))
            elif 0x7C <= op <= 0x8A:
                b = self.pop(I64); a = self.top()
                k = op - 0x7C
                if k in BIN_I:
                    E("%s=%s%s%s;" % (a, a, BIN_I[k], b))
                elif k == 3:
                    E("%s=div_s64(%s,%s);" % (a, a, b))
                elif k == 4:
                    E("if(!%s)TRAP(%d);%s=%s/%s;" % (b, self.fidx, a, a, b))
                elif k == 5:
                    E("%s=rem_s64(%s,%s);" % (a, a, b))
                elif k == 6:
                    E("if(!%s)TRAP(%d);%s=%s%%%s;" % (b, self.fidx, a, a, b))
                elif k == 10:
                    E("%s=%s<<(%s&63);" % (a, a, b))
                elif k == 11:
                    E("%s=(u64)((int64_t)%s>>(%s&63));" % (a, a, b))
                elif k == 12:
                    E("%s=%s>>(%s&63);" % (a, a, b))
                elif k == 13:
                    E("%s=rotl64(%s,%s);" % (a, a, b))
                elif k == 14:
                    E("%s=rotr64(%s,%s);" % (a, a, b))
            # ---- float arith ----
            elif 0x8B <= op <= 0x91:
                a = self.top()
                fn = ["fabsf", "-", "ceilf", "floorf", "truncf", "nearbyintf", "sqrtf"][op - 0x8B]
                E("%s=%s(%s);" % (a, fn, a))
            elif 0x92 <= op <= 0x98:
                b = self.pop(F32); a = self.top()
                k = op - 0x92
                if k < 4:
                    E("%s=%s%s%s;" % (a, a, "+-*/"[k], b))
                else:
                    E("%s=%s(%s,%s);" % (a, ["fminf", "fmaxf", "copysignf"][k - 4], a, b))
            elif 0x99 <= op <= 0x9F:
                a = self.top()
                fn = ["fabs", "-", "ceil", "floor", "trunc", "nearbyint", "sqrt"][op - 0x99]
                E("%s=%s(%s);" % (a, fn, a))
            elif 0xA0 <= op <= 0xA6:
                b = self.pop(F64); a = self.top()
                k = op - 0xA0
                if k < 4:
                    E("%s=%s%s%s;" % (a, a, "+-*/"[k], b))
                else:
                    E("%s=%s(%s,%s);" % (a, ["fmin", "fmax", "copysign"][k - 4], a, b))
            # ---- conversions ----
            elif op == 0xA7:
                a = self.pop(I64); E("%s=(u32)%s;" % (self.push(I32), a))
            elif op in (0xA8, 0xA9, 0xAA, 0xAB):
                src = F32 if op in (0xA8, 0xA9) else F64
                a = self.pop(src)
                ty = "int32_t" if op in (0xA8, 0xAA) else "u32"
                E("%s=(u32)(%s)%s;" % (self.push(I32), ty, a))
            elif op == 0xAC:
                a = self.pop(I32); E("%s=(u64)(int64_t)(int32_t)%s;" % (self.push(I64), a))
            elif op == 0xAD:
                a = self.pop(I32); E("%s=(u64)%s;" % (self.push(I64), a))
            elif op in (0xAE, 0xAF, 0xB0, 0xB1):
                src = F32 if op in (0xAE, 0xAF) else F64
                a = self.pop(src)
                ty = "int64_t" if op in (0xAE, 0xB0) else "u64"
                E("%s=(u64)(%s)%s;" % (self.push(I64), ty, a))
            elif op in (0xB2, 0xB3, 0xB4, 0xB5, 0xB6):
                src = {0xB2: (I32, "(int32_t)"), 0xB3: (I32, ""), 0xB4: (I64, "(int64_t)"), 0xB5: (I64, ""), 0xB6: (F64, "")}[op]
                a = self.pop(src[0]); E("%s=(f32)%s%s;" % (self.push(F32), src[1], a))
            elif op in (0xB7, 0xB8, 0xB9, 0xBA, 0xBB):
                src = {0xB7: (I32, "(int32_t)"), 0xB8: (I32, ""), 0xB9: (I64, "(int64_t)"), 0xBA: (I64, ""), 0xBB: (F32, "")}[op]
                a = self.pop(src[0]); E("%s=(f64)%s%s;" % (self.push(F64), src[1], a))
            elif op == 0xBC:
                a = self.pop(F32); E("%s=bits_f32(%s);" % (self.push(I32), a))
            elif op == 0xBD:
                a = self.pop(F64); E("%s=bits_f64(%s);" % (self.push(I64), a))
            elif op == 0xBE:
                a = self.pop(I32); E("%s=f32_bits(%s);" % (self.push(F32), a))
            elif op == 0xBF:
                a = self.pop(I64); E("%s=f64_bits(%s);" % (self.push(F64), a))
            elif op == 0xC0:
                a = self.top(); E("%s=(u32)(int32_t)(int8_t)%s;" % (a, a))
            elif op == 0xC1:
                a = self.top(); E("%s=(u32)(int32_t)(int16_t)%s;" % (a, a))
            elif op == 0xC2:
                a = self.top(); E("%s=(u64)(int64_t)(int8_t)%s;" % (a, a))
            elif op == 0xC3:
                a = self.top(); E("%s=(u64)(int64_t)(int16_t)%s;" % (a, a))
            elif op == 0xC4:
                a = self.top(); E("%s=(u64)(int64_t)(int32_t)%s;" % (a, a))
            elif op == 0xFC:
                sub = r.u()
                if sub <= 7:
                    src = F32 if sub in (0, 1, 4, 5) else F64
                    dst = I32 if sub < 4 else I64
                    fn = ["sat_s32", "sat_u32", "sat_s32", "sat_u32", "sat_s64", "sat_u64", "sat_s64", "sat_u64"][sub]
                    a = self.pop(src)
                    E("%s=%s((f64)%s);" % (self.push(dst), fn, a))
                elif sub == 10:
                    r.byte(); r.byte()
                    n = self.pop(I32); s = self.pop(I32); d = self.pop(I32)
                    E("memmove(mem+%s,mem+%s,%s);" % (d, s, n))
                elif sub == 11:
                    r.byte()
                    n = self.pop(I32); v = self.pop(I32); d = self.pop(I32)
                    E("memset(mem+%s,(int)%s,%s);" % (d, v, n))
                else:
                    raise NotImplementedError("0xFC %d" % sub)
            else:
                raise NotImplementedError("opcode 0x%02x in f%d" % (op, self.fidx))
        assert not ctrl, "unbalanced control in f%d" % self.fidx
        # assemble
        hdr = [self.signature() + "{"]
        if self.fidx in HOOKS:
            hdr.append("if(w2c_hook_fn)w2c_hook_fn(%d,%s);" % (
                self.fidx, ",".join(["(u64)l%d" % i for i in range(min(len(params), 4))] + ["0"] * (4 - min(len(params), 4)))))
        for i in range(len(params), len(ltypes)):
            hdr.append("%s l%d=0;" % (CT[ltypes[i]], i))
        for (d, t) in sorted(self.maxdepth):
            hdr.append("%s s%d%s;" % (CT[t], d, SUF[t]))
        return "\n".join(hdr + self.out + ["}"])

    def skip_imm(self, op, r):
        if op == 0xFC:
            sub = r.u()
            if sub == 10:
                r.byte(); r.byte()
            elif sub == 11:
                r.byte()
        # all remaining numeric ops have no immediates

    def signature(self):
        p, q = self.func_type(self.fidx)
        args = ",".join("%s l%d" % (CT[t], i) for i, t in enumerate(p)) or "void"
        return "%s fn%d(%s)" % (rettype(q), self.fidx, args)

    def emit_return(self, results):
        n = len(results)
        if n == 0:
            self.emit("return;")
        elif n == 1:
            self.emit("return %s;" % self.sv(len(self.stack) - 1, results[0]))
        else:
            base = len(self.stack) - n
            vals = ",".join(self.sv(base + i, t) for i, t in enumerate(results))
            self.emit("return (%s){%s};" % (rettype(results), vals))

    def emit_call(self, target, ty):
        p, q = ty
        args = []
        for t in reversed(p):
            args.append(self.pop(t))
        args.reverse()
        call = "%s(%s)" % (target, ",".join(args))
        if len(q) == 0:
            self.emit(call + ";")
        elif len(q) == 1:
            self.emit("%s=%s;" % (self.push(q[0]), call))
        else:
            self.emit("{%s r_=%s;" % (rettype(q), call))
            for i, t in enumerate(q):
                self.emit("%s=r_.v%d;" % (self.push(t), i))
            self.emit("}")


def rettype(q):
    if len(q) == 0:
        return "void"
    if len(q) == 1:
        return CT[q[0]]
    return "ret_" + "".join(SUF[t] for t in q)


HOOKS = set(int(x) for x in os.environ.get("W2C_HOOKS", "").split(",") if x)

RUNTIME_H = r'''
#include <stdint.h>
#include <string.h>
#include <math.h>
typedef uint32_t u32; typedef uint64_t u64; typedef float f32; typedef double f64;
extern uint8_t* mem; extern u32 mem_pages;
u32 mem_grow(u32 delta);
void* tbl_get(u32 idx);
extern void (*w2c_hook_fn)(int fn, u64 a0, u64 a1, u64 a2, u64 a3);
void wasm_trap(int fn) __attribute__((noreturn));
#define TRAP(f) wasm_trap(f)
static inline u32 clz32(u32 x){return x?__builtin_clz(x):32;}
static inline u32 ctz32(u32 x){return x?__builtin_ctz(x):32;}
static inline u32 popcnt32(u32 x){return __builtin_popcount(x);}
static inline u64 clz64(u64 x){return x?__builtin_clzll(x):64;}
static inline u64 ctz64(u64 x){return x?__builtin_ctzll(x):64;}
static inline u64 popcnt64(u64 x){return __builtin_popcountll(x);}
static inline u32 rotl32(u32 x,u32 n){n&=31;return (x<<n)|(x>>((32-n)&31));}
static inline u32 rotr32(u32 x,u32 n){n&=31;return (x>>n)|(x<<((32-n)&31));}
static inline u64 rotl64(u64 x,u64 n){n&=63;return (x<<n)|(x>>((64-n)&63));}
static inline u64 rotr64(u64 x,u64 n){n&=63;return (x>>n)|(x<<((64-n)&63));}
static inline u32 div_s32(u32 a,u32 b){if(!b||(a==0x80000000u&&b==0xFFFFFFFFu))TRAP(-1);return (u32)((int32_t)a/(int32_t)b);}
static inline u32 rem_s32(u32 a,u32 b){if(!b)TRAP(-1);if(b==0xFFFFFFFFu)return 0;return (u32)((int32_t)a%(int32_t)b);}
static inline u64 div_s64(u64 a,u64 b){if(!b||(a==0x8000000000000000ull&&b==~0ull))TRAP(-1);return (u64)((int64_t)a/(int64_t)b);}
static inline u64 rem_s64(u64 a,u64 b){if(!b)TRAP(-1);if(b==~0ull)return 0;return (u64)((int64_t)a%(int64_t)b);}
static inline f32 f32_bits(u32 b){f32 f;memcpy(&f,&b,4);return f;}
static inline f64 f64_bits(u64 b){f64 f;memcpy(&f,&b,8);return f;}
static inline u32 bits_f32(f32 f){u32 b;memcpy(&b,&f,4);return b;}
static inline u64 bits_f64(f64 f){u64 b;memcpy(&b,&f,8);return b;}
static inline u32 sat_s32(f64 x){if(x!=x)return 0;if(x<=-2147483648.0)return 0x80000000u;if(x>=2147483647.0)return 0x7FFFFFFFu;return (u32)(int32_t)x;}
static inline u32 sat_u32(f64 x){if(x!=x||x<=0.0)return 0;if(x>=4294967295.0)return 0xFFFFFFFFu;return (u32)x;}
static inline u64 sat_s64(f64 x){if(x!=x)return 0;if(x<=-9223372036854775808.0)return 0x8000000000000000ull;if(x>=9223372036854775807.0)return 0x7FFFFFFFFFFFFFFFull;return (u64)(int64_t)x;}
static inline u64 sat_u64(f64 x){if(x!=x||x<=0.0)return 0;if(x>=18446744073709551615.0)return ~0ull;return (u64)x;}
'''

RUNTIME_C = r'''
#include <stdio.h>
#include <stdlib.h>
#include <setjmp.h>
#include <sys/mman.h>
void (*w2c_hook_fn)(int,u64,u64,u64,u64);
void w2c_set_hook(void (*f)(int,u64,u64,u64,u64)){w2c_hook_fn=f;}
uint8_t* mem; u32 mem_pages; static u32 mem_max_pages = 65536;
static jmp_buf trap_jmp; static int trap_armed; static char trap_msg[512];
void wasm_trap(int fn){
  if(!trap_msg[0]) snprintf(trap_msg,sizeof trap_msg,"wasm trap in f%d",fn);
  if(trap_armed) longjmp(trap_jmp,1);
  fprintf(stderr,"%s\n",trap_msg); abort();
}
u32 mem_grow(u32 delta){
  u32 old=mem_pages; if((u64)old+delta>mem_max_pages) return 0xFFFFFFFFu;
  mem_pages+=delta; return old;
}
'''


def main():
    src, outdir = sys.argv[1], sys.argv[2]
    nchunks = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    data = open(src, "rb").read()
    m = parse(data)
    m.data = data
    m.imports_f = [d for (_, _, k, d) in m.imports if k == 0]
    nimp = len(m.imports_f)
    os.makedirs(outdir, exist_ok=True)
    nfunc = nimp + len(m.func_types)

    # multi-value return structs
    rts = set()
    for p, q in m.types:
        if len(q) > 1:
            rts.add(tuple(q))
    decl = [RUNTIME_H]
    for q in sorted(rts):
        decl.append("typedef struct{%s}%s;" % ("".join("%s v%d;" % (CT[t], i) for i, t in enumerate(q)), rettype(q)))
    for i, (t, mut, init) in enumerate(m.globals):
        decl.append("extern %s g%d;" % (CT[t], i))
    for f in range(nfunc):
        ty = m.types[m.imports_f[f]] if f < nimp else m.types[m.func_types[f - nimp]]
        p, q = ty
        decl.append("%s fn%d(%s);" % (rettype(q), f, ",".join(CT[t] for t in p) or "void"))
    open(os.path.join(outdir, "w2c.h"), "w").write("\n".join(decl) + "\n")

    # function bodies, balanced by code size
    sizes = [(e - s, i) for i, (_, s, e) in enumerate(m.codes)]
    sizes.sort(reverse=True)
    chunks = [[] for _ in range(nchunks)]
    load = [0] * nchunks
    for sz, i in sizes:
        k = load.index(min(load))
        chunks[k].append(i)
        load[k] += sz
    for k, ch in enumerate(chunks):
        with open(os.path.join(outdir, "w2c_%02d.c" % k), "w") as fo:
            fo.write('#include "w2c.h"\n')
            for i in sorted(ch):
                fo.write(FuncGen(m, nimp + i, nimp).gen())
                fo.write("\n")

    # runtime: memory, table, globals, data, imports, exports
    rt = ['#include "w2c.h"', RUNTIME_C]
    for i, (t, mut, init) in enumerate(m.globals):
        assert init[0] in ("i32", "i64")
        rt.append("%s g%d=%dull;" % (CT[t], i, init[1] & 0xFFFFFFFFFFFFFFFF))
    # imports
    k = 0
    for (mod, nm, kind, desc) in m.imports:
        if kind != 0:
            raise NotImplementedError("non-function import")
        p, q = m.types[desc]
        args = ",".join("%s a%d" % (CT[t], i) for i, t in enumerate(p)) or "void"
        if "throw" in nm:
            body = "{u32 n=a1<sizeof trap_msg-1?a1:sizeof trap_msg-1;memcpy(trap_msg,mem+a0,n);trap_msg[n]=0;wasm_trap(-2);}"
        else:
            body = "{}"
        rt.append("%s fn%d(%s)%s /* import %s.%s */" % (rettype(q), k, args, body, mod, nm))
        k += 1
    # table
    tsize = m.tables[0][1] if m.tables else 0
    rt.append("static void* table[%d];" % max(tsize, 1))
    rt.append("void* tbl_get(u32 i){if(i>=%du||!table[i])wasm_trap(-3);return table[i];}" % tsize)
    init = ["static void w2c_init_once(void){", "static int done; if(done) return; done=1;"]
    npages, maxp = m.mems[0]
    init.append("mem=mmap(0,(size_t)65536*65536+65536,PROT_READ|PROT_WRITE,MAP_PRIVATE|MAP_ANONYMOUS|MAP_NORESERVE,-1,0);")
    init.append("if(mem==MAP_FAILED){perror(\"mmap\");abort();}")
    init.append("mem_pages=%d;" % npages)
    if maxp is not None:
        init.append("mem_max_pages=%d;" % maxp)
    for (tbl, off, funcs) in m.elems:
        assert off[0] == "i32"
        for j, f in enumerate(funcs):
            init.append("table[%d]=(void*)fn%d;" % (off[1] + j, f))
    blob = []
    for di, (off, b) in enumerate(m.datas):
        if off is None:
            continue
        assert off[0] == "i32"
        rt.append("static const uint8_t data%d[%d]={%s};" % (di, len(b), ",".join(str(x) for x in b)))
        init.append("memcpy(mem+%du,data%d,%d);" % (off[1] & 0xFFFFFFFF, di, len(b)))
    init.append("}")
    rt += init
    # exports: C-callable, trap-safe wrappers.  Multi-value results come back via out[].
    rt.append("const char* w2c_last_trap(void){return trap_msg;}")
    rt.append("uint8_t* w2c_memory(void){w2c_init_once();return mem;}")
    rt.append("u64 w2c_memory_bytes(void){return (u64)mem_pages*65536;}")
    for (nm, kind, idx) in m.exports:
        if kind != 0:
            continue
        p, q = m.types[m.imports_f[idx]] if idx < nimp else m.types[m.func_types[idx - nimp]]
        assert all(t in (I32, I64) for t in p + q), nm
        args = "".join("%s a%d," % (CT[t], i) for i, t in enumerate(p))
        call = "fn%d(%s)" % (idx, ",".join("a%d" % i for i in range(len(p))))
        body = ["int w2c_%s(%su64* out){" % (nm, args),
                "w2c_init_once();trap_msg[0]=0;trap_armed=1;",
                "if(setjmp(trap_jmp)){trap_armed=0;return 1;}"]
        if len(q) == 0:
            body.append(call + ";")
        elif len(q) == 1:
            body.append("out[0]=%s;" % call)
        else:
            body.append("{%s r_=%s;%s}" % (rettype(q), call, "".join("out[%d]=r_.v%d;" % (i, i) for i in range(len(q)))))
        body.append("trap_armed=0;return 0;}")
        rt.append("\n".join(body))
    open(os.path.join(outdir, "w2c_rt.c"), "w").write("\n".join(rt) + "\n")
    print("wasm2c: %d funcs (%d imports), %d chunks -> %s" % (nfunc, nimp, nchunks, outdir))


if __name__ == "__main__":
    sys.setrecursionlimit(10000)
    main()
