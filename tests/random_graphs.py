"""Random signal graphs for differential testing: one description, rendered two ways.

    make_graph(rng) -> RandomGraph        a random acyclic graph over the node library
    g.dsl()                               its `graph!` text (for the product's front end)
    g.walk(n_voices, frames, values, gates, stream_in)   a per-sample walk over the ORACLE's node primitives

The walk is built from the description, never from the DSL text or from anything the product computed, so a bug in
the front end / plan compiler / code generator / node kernels shows up as a difference.  TEST INFRASTRUCTURE
(uses oracle_lib through graph_walk)."""
import numpy as np

import graph_walk as W

f32 = np.float32
SR = 48000.0

WAVES = ["sine", "saw", "square", "triangle"]


class Src:
    """a stream source expression tree: ("node", i) | ("value", name) | ("lit", x) | ("input",) | (op, a, b) | (method, a, args)"""

    def __init__(self, kind, *args):
        self.kind, self.args = kind, args

    def text(self, g):
        k, a = self.kind, self.args
        if k == "node":
            return f"n{a[0]}.{g.nodes[a[0]]['out']}"
        if k == "value":
            return a[0]
        if k == "lit":
            return repr(float(f32(a[0])))
        if k == "input":
            return "audio_in"
        if k == "fb":
            raise AssertionError("a feedback leg is printed as its own `-> [N] ->` statement")
        if k == "cross":
            raise AssertionError("a cross-rate edge is printed as its own `[policy] src -> dst` statement")
        if k in "+-*/":
            return f"({a[0].text(g)} {k} {a[1].text(g)})"
        margs = ", ".join(repr(float(f32(x))) for x in a[1])
        return f"{a[0].text(g)}.{k}({margs})"

    def is_stream(self):
        k, a = self.kind, self.args
        if k in ("node", "input", "fb", "cross"):
            return True
        if k in ("value", "lit"):
            return False
        if k in "+-*/":
            return a[0].is_stream() or a[1].is_stream()
        return True

    def eval(self, outs, vals, xin):
        k, a = self.kind, self.args
        if k == "node":
            return outs[a[0]]
        if k == "value":
            return vals[a[0]]
        if k == "lit":
            return f32(a[0])
        if k == "input":
            return xin
        if k == "fb":                                        # the leg out of a feedback delay: what it output when it last ran
            return f32(outs["fb"][a[0]].output)
        if k == "cross":                                     # Up edge: this inner tick's sample; Down edge: this frame's result
            c = outs["cross"][a[0]]
            return f32(c[outs["tick"]]) if outs["dir"][a[0]] == "up" else f32(c)
        if k in "+-*/":
            x, y = a[0].eval(outs, vals, xin), a[1].eval(outs, vals, xin)
            with np.errstate(all="ignore"):
                return f32(x + y) if k == "+" else f32(x - y) if k == "-" else f32(x * y) if k == "*" else f32(x / y)
        x = a[0].eval(outs, vals, xin)
        if k == "tanh":
            return W.tanh32(x)
        if k == "abs":
            return f32(abs(x))
        if k == "clamp":
            return W.clamp32(x, a[1][0], a[1][1])
        if k == "min":
            return f32(min(x, f32(a[1][0])))
        if k == "max":
            return f32(max(x, f32(a[1][0])))
        return W.libm32({"sin": "sinf", "cos": "cosf", "exp": "expf", "sqrt": "sqrtf"}[k], x)


class RandomGraph:
    def __init__(self):
        self.values = {}          # name -> default
        self.nodes = []           # dicts: type, ctor text, out endpoint name, conns {endpoint: Src | ("value", name) | ("lit", x)}
        self.has_input = False
        self.has_gate = False
        self.output = None        # Src
        self.feedbacks = []       # (source node j, samples N, target node i, endpoint): `nJ.out -> [N] -> nI.ep`, i an ancestor of j (or j itself)
        self.rates = []           # per node: log2 of its rate factor (0 = the graph's rate)
        self.inner_log2 = 0
        self.cross = []           # cross-rate edges: {"dir": "up" | "down", "policy": str, "src": Src}

    # ---- DSL text
    def dsl(self):
        L = ["name: Fuzz;"]
        if self.has_input:
            L.append("input audio_in: stream;")
        if self.has_gate:
            L.append("input gate: event;")
        for name, d in self.values.items():
            L.append(f"input {name}: value = {float(f32(d))!r};")
        L.append("output out: stream;")
        L.append("nodes {")
        for i, nd in enumerate(self.nodes):
            rate = f" * {1 << self.rates[i]}" if self.rates and self.rates[i] else ""
            L.append(f"    n{i} = {nd['ctor']}{rate};")
        L.append("}")
        L.append("connections {")
        for i, nd in enumerate(self.nodes):
            for ep, srcs in nd["conns"].items():
                for s in srcs:
                    if s.kind == "fb":
                        j, samples, _, _ = self.feedbacks[s.args[0]]
                        L.append(f"    n{j}.{self.nodes[j]['out']} -> [{samples}] -> n{i}.{ep};")
                    elif s.kind == "cross":
                        c = self.cross[s.args[0]]
                        pol = "" if c["policy"] == "default" else f"[{c['policy']}] "
                        L.append(f"    {pol}{c['src'].text(self)} -> n{i}.{ep};")
                    else:
                        L.append(f"    {s.text(self)} -> n{i}.{ep};")
        if self.output.kind == "cross":
            c = self.cross[self.output.args[0]]
            pol = "" if c["policy"] == "default" else f"[{c['policy']}] "
            L.append(f"    {pol}{c['src'].text(self)} -> out;")
        else:
            L.append(f"    {self.output.text(self)} -> out;")
        L.append("}")
        return "\n".join(L)

    # ---- oracle walk
    def walk(self, n_voices, frames, values, gates, stream_in):
        def make(v):
            objs = [nd["make"]() for nd in self.nodes]
            fbs = [W.delay(float(samples), 0.0, SR) for (_, samples, _, _) in self.feedbacks]     # Delay::new(N as f32, 0.0)

            def gate_fn(s):
                for nd, o in zip(self.nodes, objs):
                    if nd["type"] == "adsr" and "gate" in nd["conns"]:
                        o.on_gate(s)

            import ctypes
            import oracle_lib as O
            L = O.lib()
            N = 1 << self.inner_log2
            res = []                                          # one resampler per cross-rate edge, like the reference's per-edge fields
            for c in self.cross:
                pol = "sinc" if c["policy"] == "default" else c["policy"]
                if c["dir"] == "up":
                    if pol == "sinc":
                        o = O.SincUp(); L.orc_sinc_up_new(ctypes.byref(o), ctypes.c_int(N)); fn = lambda x, out, o=o: L.orc_sinc_up_process(ctypes.byref(o), float(x), out)
                    elif pol == "sinc_iir":
                        o = O.IirResampler(); L.orc_iir_up_new(ctypes.byref(o), ctypes.c_int(N)); fn = lambda x, out, o=o: L.orc_iir_up_process(ctypes.byref(o), float(x), out)
                    elif pol == "linear":
                        o = O.LinearUp(); L.orc_linear_up_new(ctypes.byref(o), ctypes.c_int(N)); fn = lambda x, out, o=o: L.orc_linear_up_process(ctypes.byref(o), float(x), out)
                    else:
                        fn = lambda x, out: L.orc_latch_up_process(N, float(x), out)
                else:
                    if pol == "sinc":
                        o = O.SincDown(); L.orc_sinc_down_new(ctypes.byref(o), ctypes.c_int(N)); fn = lambda xs, o=o: L.orc_sinc_down_process(ctypes.byref(o), xs)
                    elif pol == "sinc_iir":
                        o = O.IirResampler(); L.orc_iir_down_new(ctypes.byref(o), ctypes.c_int(N)); fn = lambda xs, o=o: L.orc_iir_down_process(ctypes.byref(o), xs)
                    elif pol == "linear":
                        fn = lambda xs: L.orc_linear_down_process(N, xs)
                    else:
                        fn = lambda xs: L.orc_latch_down_process(N, xs)
                res.append(fn)

            def run_node(i, outs, vals, xin):
                nd, o = self.nodes[i], objs[i]
                ins = {}
                for ep, srcs in nd["conns"].items():
                    if ep == "gate":
                        continue
                    if nd["ep_kind"][ep] == "value":
                        ins[ep] = srcs[-1].eval(outs, vals, xin)              # last write wins
                    else:
                        acc = srcs[0].eval(outs, vals, xin)
                        for s in srcs[1:]:
                            acc = f32(acc + s.eval(outs, vals, xin))          # fan-in: connect the first, accumulate the rest
                        ins[ep] = acc
                outs[i] = nd["run"](o, ins)
                for k, (j, _, _, _) in enumerate(self.feedbacks):              # the hidden delay runs once its source has
                    if j == i:
                        fbs[k].set(input=outs[i]); fbs[k].process()

            def frame(vals, xin):
                outs = {k: None for k in range(len(self.nodes))}
                outs["fb"] = fbs
                if not self.inner_log2:
                    for i in range(len(self.nodes)):
                        run_node(i, outs, vals, xin)
                    return self.output.eval(outs, vals, xin)
                # the reference's multirate frame body (codegen/emit_frame.rs:114-176)
                outs["cross"] = [None] * len(self.cross); outs["dir"] = [c["dir"] for c in self.cross]; outs["tick"] = 0
                inner = [i for i in range(len(self.nodes)) if self.rates[i]]
                pre = [i for i in range(len(self.nodes)) if not self.rates[i] and i < inner[0]]
                post = [i for i in range(len(self.nodes)) if not self.rates[i] and i > inner[0]]
                for i in pre:
                    run_node(i, outs, vals, xin)
                for k, c in enumerate(self.cross):
                    if c["dir"] == "up":
                        buf = (ctypes.c_float * N)(); res[k](c["src"].eval(outs, vals, xin), buf); outs["cross"][k] = buf
                acc = {k: (ctypes.c_float * N)() for k, c in enumerate(self.cross) if c["dir"] == "down"}
                for tick in range(N):
                    outs["tick"] = tick
                    for i in inner:
                        run_node(i, outs, vals, xin)
                    for k in acc:
                        acc[k][tick] = self.cross[k]["src"].eval(outs, vals, xin)
                for k in acc:
                    outs["cross"][k] = f32(res[k](acc[k]))
                for i in post:
                    run_node(i, outs, vals, xin)
                return self.output.eval(outs, vals, xin)
            return frame, gate_fn
        return W.render(make, n_voices, frames, values, gates, stream_in)


def _stateless(fn):
    return lambda o, ins: fn(ins)


# extended mode: value endpoints the base generator leaves at their constructor values, by kind of quantity
EXT_ENDPOINTS = {
    "polyblep": [("frequency", "freq")], "osc": [("frequency", "freq")],
    "adsr": [("attack", "time"), ("decay", "time"), ("release", "time")],
    "lp18": [("cutoff", "freq"), ("fmod", "unit")], "iirlp": [("cutoff", "freq")],
    "delay": [("delay_samples", "samples"), ("feedback", "unit")],
}
EXT_RANGES = {"freq": (100.0, 3000.0), "time": (0.0005, 0.003), "unit": (0.05, 0.6), "samples": (0.0, 35.0)}


def make_graph(rng, n_nodes=None, allow_delay=True, feedback=False, multirate=False, extended=False):
    """multirate: nodes [a, b) run at 2^L x the graph's rate (`* N`); every stream edge between the two rates goes through
    its own resampler of a random connection policy (one source per cross-rate endpoint: no fan-in across rates).
    extended: additionally drives the value endpoints of EXT_ENDPOINTS (oscillator frequencies, envelope times, filter
    cutoffs, delay time / feedback) from dedicated graph value inputs or literals.  The extra choices come from a second
    generator seeded from `rng`'s state, so a seed yields the SAME base graph with and without the flag."""
    g = RandomGraph()
    rng2 = np.random.default_rng(int(rng.bit_generator.state["state"]["state"]) & 0xFFFFFFFF) if extended else None
    ext_values = {}
    n_nodes = n_nodes or int(rng.integers(3, 9))
    n_values = int(rng.integers(0, 4))
    for k in range(n_values):
        g.values[f"p{k}"] = float(f32(rng.uniform(0.1, 0.9)))
    g.has_input = bool(rng.random() < 0.4)
    g.rates = [0] * n_nodes
    lo = hi = 0
    if multirate:
        g.inner_log2 = int(rng.integers(1, 4))
        lo = int(rng.integers(0, n_nodes - 1)) if (g.has_input or n_nodes > 3) else 0
        if lo == 0 and not g.has_input:
            lo = 0                                              # an inner-rate source node opens the graph
        hi = int(rng.integers(lo + 1, n_nodes + 1))
        for i in range(lo, hi):
            g.rates[i] = g.inner_log2

    def stream_leaf(i, ctx=None):
        """a stream-kind leaf available to node i (earlier node outputs of ITS rate, the graph's stream input at the outer
        rate); ctx = (upper node bound, rate) looks at another rate's nodes instead (the source side of a cross-rate edge)"""
        bound, rate = ctx if ctx is not None else (i, g.rates[i])
        choices = [Src("node", j, ) for j in range(bound) if g.rates[j] == rate]
        if g.has_input and rate == 0:
            choices.append(Src("input"))
        if not choices:
            return None
        return choices[int(rng.integers(len(choices)))]

    def value_leaf():
        if g.values and rng.random() < 0.6:
            return Src("value", list(g.values)[int(rng.integers(len(g.values)))])
        return Src("lit", float(f32(rng.uniform(0.05, 1.5))))

    def expr(i, depth=0, ctx=None):
        """a stream-kind expression for an input of node i (or None when nothing is available yet)"""
        s = stream_leaf(i, ctx)
        if s is None:
            return None
        r = rng.random()
        if depth >= 2 or r < 0.45:
            return s
        if r < 0.75:
            op = "+-*/"[int(rng.integers(4))]
            other = expr(i, depth + 1, ctx) if rng.random() < 0.5 else value_leaf()
            if other is None:
                other = value_leaf()
            if op == "/":                                   # keep divisors away from zero: |x| + c
                other = Src("+", Src("abs", other, ()), Src("lit", 0.5)) if other.is_stream() else Src("lit", 0.75)
            a, b = (s, other) if rng.random() < 0.7 else (other, s)
            if op == "/" and a is other:
                a, b = s, other
            return Src(op, a, b)
        m = ["tanh", "abs", "clamp", "min", "max", "sin", "cos", "sqrt"][int(rng.integers(8))]
        inner = expr(i, depth + 1, ctx) or s
        if m == "clamp":
            return Src("clamp", inner, (-0.5, 0.6))
        if m in ("min", "max"):
            return Src(m, inner, (0.25,))
        if m == "sqrt":
            return Src("sqrt", Src("abs", inner, ()), ())
        return Src(m, inner, ())

    def maybe(i, p=0.7):
        return expr(i) if rng.random() < p else None

    for i in range(n_nodes):
        kinds = ["polyblep", "osc", "tpt", "lp18", "iirlp", "adsr", "fmop", "gain", "addvalue", "mixer", "crossfade", "hardclip"]
        if allow_delay:
            kinds.append("delay")
        t = kinds[int(rng.integers(len(kinds)))]
        if i == 0 and not g.has_input and t not in ("polyblep", "osc", "adsr", "fmop"):
            t = "polyblep"                                   # the first node must be a source
        nd = {"type": t, "conns": {}, "out": "output", "ep_kind": {}}
        sr_i = SR * (1 << g.rates[i])                        # inner-rate nodes: set_sample_rate(sr * N)

        def conn(ep, kind, src):
            if multirate and kind == "stream" and ep not in nd["conns"] and (src is None or rng.random() < 0.35):
                # the source side of a cross-rate edge: an expression at the OTHER rate (outer nodes before the section for an
                # inner destination; the section's nodes for an outer destination after it), through its own resampler
                other = (lo, 0) if g.rates[i] else ((hi, g.inner_log2) if i >= hi else None)
                xs = expr(i, 1, other) if other is not None else None
                if xs is not None and xs.is_stream():
                    g.cross.append({"dir": "up" if g.rates[i] else "down", "src": xs,
                                    "policy": ["default", "sinc", "sinc_iir", "linear", "latch"][int(rng.integers(5))]})
                    src = Src("cross", len(g.cross) - 1)
            if src is None:
                return
            if multirate and ep in nd["conns"] and (src.kind == "cross" or nd["conns"][ep][0].kind == "cross"):
                return                                          # no fan-in across rates
            nd["ep_kind"][ep] = kind
            nd["conns"].setdefault(ep, []).append(src)

        def value_conn(ep, p=0.5, scale=1.0):
            if rng.random() < p:
                if g.values and rng.random() < 0.7:
                    conn(ep, "value", Src("value", list(g.values)[int(rng.integers(len(g.values)))]))
                else:
                    conn(ep, "value", Src("lit", float(f32(rng.uniform(0.1, 0.9) * scale))))

        if t == "polyblep":
            w = WAVES[int(rng.integers(4))]; f = float(f32(rng.uniform(50, 3000))); a = float(f32(rng.uniform(0.2, 1.0)))
            nd["ctor"] = f"PolyBlepOscillator::{w}({f!r}, {a!r})"
            nd["make"] = lambda w=w, f=f, a=a, sr_i=sr_i: W.polyblep(w, f, a, sr_i)
            conn("frequency_mod", "stream", maybe(i, 0.5)); conn("phase_mod", "stream", maybe(i, 0.4))
            value_conn("amplitude", 0.3)
            nd["run"] = lambda o, ins: (o.set(frequency_mod=ins.get("frequency_mod", o.frequency_mod), phase_mod=ins.get("phase_mod", o.phase_mod),
                                              amplitude=ins.get("amplitude", o.amplitude)), o.process())[1]
        elif t == "osc":
            w = ["sine", "square", "saw"][int(rng.integers(3))]; f = float(f32(rng.uniform(50, 3000))); a = float(f32(rng.uniform(0.2, 1.0)))
            nd["ctor"] = f"Oscillator::{w}({f!r}, {a!r})"
            nd["make"] = lambda w=w, f=f, a=a, sr_i=sr_i: W.naive_osc(w, f, a, sr_i)
            conn("frequency_mod", "stream", maybe(i, 0.5)); value_conn("amplitude", 0.3)
            nd["run"] = lambda o, ins: (o.set(frequency_mod=ins.get("frequency_mod", o.frequency_mod), amplitude=ins.get("amplitude", o.amplitude)), o.process())[1]
        elif t == "tpt":
            c = float(f32(rng.uniform(200, 8000))); q = float(f32(rng.uniform(0.3, 4.0)))
            nd["ctor"] = f"TptFilter::new({c!r}, {q!r})"
            nd["make"] = lambda c=c, q=q, sr_i=sr_i: W.tpt(c, q, sr_i)
            conn("input", "stream", expr(i)); conn("f_mod", "stream", maybe(i, 0.3))
            if rng.random() < 0.4:
                e = expr(i)
                if e is not None:                            # cutoff from a stream: offset into an audible range
                    conn("cutoff", "stream", Src("+", Src("*", Src("abs", e, ()), Src("lit", 2000.0)), Src("lit", 300.0)))
            value_conn("q", 0.3, 2.0)
            nd["run"] = lambda o, ins: (o.set(input=ins.get("input", o.input), cutoff=ins.get("cutoff", o.cutoff), f_mod=ins.get("f_mod", o.f_mod),
                                              q=ins.get("q", o.q)), o.process())[1]
        elif t == "lp18":
            c = float(f32(rng.uniform(200, 8000))); r = float(f32(rng.uniform(0.0, 0.9)))
            nd["ctor"] = f"LP18Filter::new({c!r}, {r!r})"
            nd["make"] = lambda c=c, r=r, sr_i=sr_i: W.lp18(c, r, sr_i)
            conn("input", "stream", expr(i)); value_conn("resonance", 0.3)
            nd["run"] = lambda o, ins: (o.set(input=ins.get("input", o.input), resonance=ins.get("resonance", o.resonance)), o.process())[1]
        elif t == "iirlp":
            c = float(f32(rng.uniform(200, 8000))); q = float(f32(rng.uniform(0.3, 4.0)))
            nd["ctor"] = f"IirLowpass::new({c!r}, {q!r})"
            nd["make"] = lambda c=c, q=q, sr_i=sr_i: W.iirlp(c, q, sr_i)
            conn("input", "stream", expr(i)); value_conn("q", 0.3, 2.0)
            nd["run"] = lambda o, ins: (o.set(input=ins.get("input", o.input), q=ins.get("q", o.q)), o.process())[1]
        elif t == "adsr":
            a_, d_, s_, r_ = (float(f32(x)) for x in (rng.uniform(0.0005, 0.002), rng.uniform(0.0005, 0.003), rng.uniform(0.2, 0.9), rng.uniform(0.0005, 0.002)))
            nd["ctor"] = f"AdsrEnvelope::new({a_!r}, {d_!r}, {s_!r}, {r_!r})"
            nd["make"] = lambda a_=a_, d_=d_, s_=s_, r_=r_, sr_i=sr_i: W.adsr(a_, d_, s_, r_, sr_i)
            if rng.random() < 0.8:
                g.has_gate = True
                nd["conns"]["gate"] = [None]
            value_conn("sustain", 0.4)
            nd["run"] = lambda o, ins: (o.set(sustain=ins.get("sustain", o.sustain)), o.process())[1]
        elif t == "fmop":
            nd["ctor"] = "FmOperator::new()"
            nd["make"] = lambda sr_i=sr_i: W.fmop(sr_i)
            conn("phase_mod", "stream", maybe(i, 0.6)); conn("envelope", "stream", maybe(i, 0.5))
            value_conn("ratio", 0.4, 3.0); value_conn("feedback", 0.4, 0.5); value_conn("level", 0.3)
            conn("base_freq", "value", Src("lit", float(f32(rng.uniform(80, 900)))))
            nd["run"] = lambda o, ins: (o.set(**{k: ins[k] for k in ("phase_mod", "envelope", "ratio", "feedback", "level", "base_freq") if k in ins}), o.process())[1]
        elif t == "gain":
            gv = float(f32(rng.uniform(0.2, 1.5)))
            nd["ctor"] = f"Gain::new({gv!r})"
            nd["make"] = lambda: None
            conn("input", "stream", expr(i))
            if rng.random() < 0.5:
                conn("gain", "stream", maybe(i, 1.0) if rng.random() < 0.5 else value_leaf())
            nd["run"] = _stateless(lambda ins, gv=gv: f32(ins.get("input", f32(0)) * ins.get("gain", f32(gv))))
        elif t == "addvalue":
            av = float(f32(rng.uniform(-0.5, 0.5)))
            nd["ctor"] = f"AddValue::new({av!r})"
            nd["make"] = lambda: None
            conn("input", "stream", expr(i)); value_conn("value", 0.5)
            nd["run"] = _stateless(lambda ins, av=av: f32(ins.get("input", f32(0)) + ins.get("value", f32(av))))
        elif t == "mixer":
            nd["ctor"] = "Mixer::new()"
            nd["make"] = lambda: None
            conn("input_a", "stream", expr(i)); conn("input_b", "stream", maybe(i, 0.8))
            if rng.random() < 0.3:
                conn("input_a", "stream", expr(i))           # fan-in
            nd["run"] = _stateless(lambda ins: f32(ins.get("input_a", f32(0)) + ins.get("input_b", f32(0))))
        elif t == "crossfade":
            nd["ctor"] = "Crossfade::new()"
            nd["make"] = lambda: None
            nd["out"] = "output_a" if rng.random() < 0.5 else "output_b"
            conn("input", "stream", expr(i)); value_conn("mix", 0.8)
            which = nd["out"]
            def run_xf(ins, which=which):
                mix = W.clamp32(ins.get("mix", f32(0)), 0.0, 1.0)
                x = ins.get("input", f32(0))
                return f32(x * f32(f32(1.0) - mix)) if which == "output_a" else f32(x * mix)
            nd["run"] = _stateless(run_xf)
        elif t == "hardclip":
            nd["ctor"] = "HardClip::new()"
            nd["make"] = lambda: None
            conn("input", "stream", expr(i))
            nd["run"] = _stateless(lambda ins: W.hardclip(ins.get("input", f32(0))))
        else:  # delay: constant whole-sample or fractional time, own feedback
            ds = float(f32(int(rng.integers(0, 40)))) if rng.random() < 0.6 else float(f32(rng.uniform(0.0, 30.0)))
            fb = float(f32(rng.uniform(0.0, 0.7)))
            nd["ctor"] = f"Delay::new({ds!r}, {fb!r})"
            nd["make"] = lambda ds=ds, fb=fb, sr_i=sr_i: W.delay(ds, fb, sr_i)
            conn("input", "stream", expr(i))
            nd["run"] = lambda o, ins: (o.set(input=ins.get("input", o.input)), o.process())[1]
        if extended:
            eps = []
            for ep, vk in EXT_ENDPOINTS.get(t, []):
                if ep in nd["conns"] or rng2.random() < 0.45:
                    continue
                lo_v, hi_v = EXT_RANGES[vk]
                if rng2.random() < 0.7:                      # a dedicated graph value input per kind of quantity (shared by the nodes)
                    name = "x" + vk
                    if name not in ext_values:               # merged into g.values at the end: the base choices must not see them
                        ext_values[name] = float(f32(rng2.uniform(lo_v, hi_v)))
                    src = Src("value", name)
                else:
                    src = Src("lit", float(f32(rng2.uniform(lo_v, hi_v))))
                nd["ep_kind"][ep] = "value"
                nd["conns"][ep] = [src]
                eps.append(ep)
            if eps:
                base_run = nd["run"]
                nd["run"] = lambda o, ins, base_run=base_run, eps=eps: (o.set(**{k: ins[k] for k in eps if k in ins}), base_run(o, ins))[1]
        # drop inputs that could not be connected (nothing available yet)
        nd["conns"] = {ep: [s for s in srcs if s is not None or ep == "gate"] for ep, srcs in nd["conns"].items()}
        nd["conns"] = {ep: srcs for ep, srcs in nd["conns"].items() if srcs}
        g.nodes.append(nd)
    g.values.update(ext_values)
    # feedback vias: `nJ.out -> [N] -> nI.<stream input>` with I an ancestor of J (or J itself), so the edge closes a cycle
    if feedback:
        def ancestors(j):
            seen, stack = set(), [j]
            while stack:
                k = stack.pop()
                for srcs in g.nodes[k]["conns"].values():
                    for s0 in srcs:
                        st = [s0]
                        while st:
                            e = st.pop()
                            if e is None:
                                continue
                            if e.kind == "node":
                                if e.args[0] not in seen:
                                    seen.add(e.args[0]); stack.append(e.args[0])
                            elif e.kind in "+-*/":
                                st += [e.args[0], e.args[1]]
                            elif e.kind not in ("value", "lit", "input", "fb"):
                                st.append(e.args[0])
            return seen
        stream_eps = {"polyblep": ["frequency_mod", "phase_mod"], "osc": ["frequency_mod"], "tpt": ["input", "f_mod"], "lp18": ["input"],
                      "iirlp": ["input"], "fmop": ["phase_mod"], "gain": ["input"], "addvalue": ["input"], "mixer": ["input_a", "input_b"],
                      "crossfade": ["input"], "hardclip": ["input"], "delay": ["input"]}
        for _ in range(int(rng.integers(1, 3))):
            j = int(rng.integers(0, n_nodes))
            cands = sorted(ancestors(j) | {j})
            cands = [i for i in cands if g.nodes[i]["type"] in stream_eps]
            if not cands:
                continue
            i = cands[int(rng.integers(len(cands)))]
            eps = stream_eps[g.nodes[i]["type"]]
            ep = eps[int(rng.integers(len(eps)))]
            samples = int(rng.integers(1, 6)) if rng.random() < 0.7 else int(rng.integers(6, 70))
            g.feedbacks.append((j, samples, i, ep))
            g.nodes[i]["ep_kind"][ep] = "stream"
            g.nodes[i]["conns"].setdefault(ep, []).append(Src("fb", len(g.feedbacks) - 1))
    # gate connections print as `gate -> nI.gate`
    for nd in g.nodes:
        if "gate" in nd["conns"]:
            nd["conns"]["gate"] = [Src("value", "gate")]
            nd["ep_kind"]["gate"] = "event"
    # the output: the last node, possibly combined with another one
    out = Src("node", n_nodes - 1)
    if multirate:
        if g.rates[n_nodes - 1]:                             # the graph ends inside the oversampled section: a Down edge to the output
            g.cross.append({"dir": "down", "src": out, "policy": ["default", "sinc", "sinc_iir", "linear", "latch"][int(rng.integers(5))]})
            out = Src("cross", len(g.cross) - 1)
    elif n_nodes > 1 and rng.random() < 0.5:
        out = Src("+", out, Src("*", Src("node", int(rng.integers(0, n_nodes - 1))), Src("lit", 0.5)))
    g.output = out
    return g
