"""oscen_b200 — B200-native, voice-batched engine for Oscen's signal-graph evaluation loop.

The product is the C-ABI shared library (include/oscen_b200.h, oscen_b200/csrc/*) — hand-written
sm_100a CUDA behind the API surface Oscen's `graph!` macro generates.  This package is the thin
Python host mirror used by the tests and bench.py; it has no CPU rendering path."""
from . import abi  # noqa: F401
from .abi import (Engine, EngineError, EVENT_GATE, GRAPH_EPIANO, GRAPH_FM_VOICE, GRAPH_SAT_1X,  # noqa: F401
                  GRAPH_SAT_4X, GRAPH_SYNTH, GRAPH_SYNTH_ADSR, GRAPH_TWIN_PEAKS, KEEP_INSTANCE_OUT,
                  KEEP_VOICE_OUT, PLAN_FUSED, PLAN_NODES)
