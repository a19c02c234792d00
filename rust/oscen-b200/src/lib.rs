//! Safe layer over `oscen-b200-sys`.  A [`Bank`] is N instances of one generated graph struct rendered together on a
//! B200; its methods are the ones `graph!` generates (`oscen-graph-compiler/src/codegen/mod.rs:1292-1392`):
//! `new` + `init` (here: [`Bank::new`]), `set_<v>` ([`Bank::set`]), `set_<v>_with_ramp`, `set_<v>_immediate`,
//! `<event>.try_push` ([`Bank::push_gate`], [`Bank::push_midi`]), `process_block`, `<out>_block` ([`Bank::out_block`]),
//! `latency_samples`, and `BlockRender::render` (`oscen-lib/src/graph/offline.rs:46-90`).
//! Like the generated struct a `Bank` is `Send` but not `Sync`: one caller thread at a time (`&mut self`).
use oscen_b200_sys as sys;
use std::ffi::{CStr, CString};
use std::os::raw::c_char;

#[derive(Debug)]
pub struct Error(pub String);
pub type Result<T> = std::result::Result<T, Error>;

/// A `graph!` description compiled to a launch plan (needs no GPU).
pub struct Graph {
    raw: *mut sys::oscen_b200_graph,
}
unsafe impl Send for Graph {}

impl Graph {
    /// `text` is the body of a `graph! { ... }` invocation (see [`b200_graph!`]); with several blocks the last is compiled.
    pub fn compile(text: &str) -> Result<Self> {
        let c = CString::new(text).map_err(|_| Error("graph text contains a NUL byte".into()))?;
        let mut raw = std::ptr::null_mut();
        let mut err = [0 as c_char; 1024];
        let rc = unsafe { sys::oscen_b200_graph_compile(c.as_ptr(), &mut raw, err.as_mut_ptr(), err.len()) };
        if rc != 0 {
            return Err(Error(unsafe { CStr::from_ptr(err.as_ptr()) }.to_string_lossy().into_owned()));
        }
        Ok(Self { raw })
    }
    pub fn input_id(&self, name: &str) -> Option<u32> {
        let c = CString::new(name).ok()?;
        let id = unsafe { sys::oscen_b200_graph_input_id(self.raw, c.as_ptr()) };
        (id >= 0).then_some(id as u32)
    }
    /// frames the generated `set_<v>()` ramps over (`ramp: N` in the input's spec), 0 = plain field
    pub fn ramp_frames(&self, input_id: u32) -> u32 {
        let mut info = std::mem::MaybeUninit::<sys::oscen_b200_input_info>::zeroed();
        let rc = unsafe { sys::oscen_b200_graph_input_info(self.raw, input_id, info.as_mut_ptr()) };
        if rc != 0 { 0 } else { unsafe { info.assume_init() }.ramp_frames }
    }
    pub fn voices_per_instance(&self) -> u32 {
        unsafe { sys::oscen_b200_graph_voices_per_instance(self.raw) }
    }
    /// lanes of the graph's stream output: 1, or 2 for `output out: stream: Frame<2>;` (oscen-lib/src/frame.rs:57-82)
    pub fn channels(&self) -> u32 {
        unsafe { sys::oscen_b200_graph_channels(self.raw) }
    }
    /// index of an `external NAME: AudioAsset;` slot
    pub fn asset_id(&self, name: &str) -> Option<u32> {
        let c = CString::new(name).ok()?;
        let id = unsafe { sys::oscen_b200_graph_asset_id(self.raw, c.as_ptr()) };
        (id >= 0).then_some(id as u32)
    }
}
impl Drop for Graph {
    fn drop(&mut self) {
        unsafe { sys::oscen_b200_graph_free(self.raw) }
    }
}

/// N instances of a compiled graph on one GPU.
pub struct Bank {
    raw: *mut sys::oscen_b200_engine,
    graph: Graph,
    instances: u32,
    vpi: u32,
    channels: u32,
    frames: u32,
}
unsafe impl Send for Bank {}

impl Bank {
    pub const MAX_BLOCK_SIZE: u32 = sys::OSCEN_B200_MAX_BLOCK; // DEFAULT_MAX_BLOCK_SIZE, oscen-lib/src/graph/types.rs:12

    /// `G::new()` + `G::init(sample_rate)` for every instance; the graph runs as one generated kernel
    pub fn new(graph: Graph, instances: u32, sample_rate: f32, device: i32) -> Result<Self> {
        let vpi = graph.voices_per_instance();
        let cfg = sys::oscen_b200_config {
            graph_kind: sys::OSCEN_B200_GRAPH_CUSTOM,
            n_voices: instances * vpi,
            voices_per_instance: vpi,
            sample_rate,
            device,
            plan: sys::OSCEN_B200_PLAN_FUSED,
            keep: sys::OSCEN_B200_KEEP_INSTANCE_OUT,
            reserved: 0,
        };
        let mut raw = std::ptr::null_mut();
        let rc = unsafe { sys::oscen_b200_create_from_graph(graph.raw, &cfg, &mut raw) };
        if rc != 0 {
            return Err(Error(unsafe { CStr::from_ptr(sys::oscen_b200_last_error(std::ptr::null())) }.to_string_lossy().into_owned()));
        }
        let channels = unsafe { sys::oscen_b200_channels(raw) };
        Ok(Self { raw, graph, instances, vpi, channels, frames: 0 })
    }
    fn check(&self, rc: i32, what: &str) -> Result<()> {
        if rc == 0 {
            Ok(())
        } else {
            let msg = unsafe { CStr::from_ptr(sys::oscen_b200_last_error(self.raw)) }.to_string_lossy().into_owned();
            Err(Error(format!("{what}: {msg}")))
        }
    }
    fn input(&self, name: &str) -> Result<u32> {
        self.graph.input_id(name).ok_or_else(|| Error(format!("graph has no value input `{name}`")))
    }
    /// the generated `set_<name>(v)`: ramps over the declared `ramp: N` frames, or assigns
    pub fn set(&mut self, name: &str, v: f32) -> Result<()> {
        let id = self.input(name)?;
        match self.graph.ramp_frames(id) {
            0 => self.set_immediate(name, v),
            n => self.check(unsafe { sys::oscen_b200_set_value_ramp(self.raw, id, v, n) }, "set_value_ramp"),
        }
    }
    pub fn set_with_ramp(&mut self, name: &str, v: f32, frames: u32) -> Result<()> {
        let id = self.input(name)?;
        self.check(unsafe { sys::oscen_b200_set_value_ramp(self.raw, id, v, frames) }, "set_value_ramp")
    }
    pub fn set_immediate(&mut self, name: &str, v: f32) -> Result<()> {
        let id = self.input(name)?;
        self.check(unsafe { sys::oscen_b200_set_value(self.raw, id, 0, self.instances * self.vpi, &v, 1) }, "set_value")
    }
    /// `gate.try_push(EventInstance { frame_offset, payload: Scalar(velocity) })` on one voice
    pub fn push_gate(&mut self, voice: u32, frame_offset: u32, velocity: f32) -> Result<()> {
        let ev = sys::oscen_b200_event { voice, frame_offset, input_id: sys::OSCEN_B200_EVENT_GATE, scalar: velocity };
        self.check(unsafe { sys::oscen_b200_push_events(self.raw, &ev, 1) }, "push_events")
    }
    /// `midi_in.try_push(raw_midi_event(bytes))` on one instance of a polyphonic graph
    pub fn push_midi(&mut self, instance: u32, frame_offset: u32, bytes: &[u8]) -> Result<()> {
        self.check(unsafe { sys::oscen_b200_push_midi(self.raw, instance, frame_offset, bytes.as_ptr(), bytes.len() as u32) }, "push_midi")
    }
    /// `graph.<slot>.load(..)` of an `external <slot>: AudioAsset;`: `interleaved` holds frames x channels f32 at the
    /// bank's sample rate; the SamplePlayers bound to the slot restart at the next block
    pub fn load_asset(&mut self, slot: &str, interleaved: &[f32], channels: u32) -> Result<()> {
        let id = self.graph.asset_id(slot).ok_or_else(|| Error(format!("graph has no external slot `{slot}`")))?;
        let ch = channels.max(1);
        let frames = (interleaved.len() as u32) / ch;
        self.check(unsafe { sys::oscen_b200_load_asset(self.raw, id, interleaved.as_ptr(), frames, ch) }, "load_asset")
    }
    /// (no reference counterpart) compile the kernel variant that reads the value inputs which hold one value for the whole
    /// bank from the constant bank; returns the ids taken as a bit mask.  Not for the audio thread: it runs NVRTC.
    pub fn specialize(&mut self) -> Result<u32> {
        let mut mask = 0u32;
        self.check(unsafe { sys::oscen_b200_specialize(self.raw, &mut mask) }, "specialize")?;
        Ok(mask)
    }
    /// `process_block(frames)`, frames <= MAX_BLOCK_SIZE
    pub fn process_block(&mut self, frames: u32) -> Result<()> {
        self.check(unsafe { sys::oscen_b200_process_block(self.raw, frames) }, "process_block")?;
        self.frames = frames;
        Ok(())
    }
    /// every instance's `<out>_block` of the last block: `[frame][instance][channel]`
    pub fn out_block(&mut self) -> Result<Vec<f32>> {
        let mut out = vec![0.0f32; (self.frames * self.instances * self.channels) as usize];
        self.check(unsafe { sys::oscen_b200_read_instance_out(self.raw, out.as_mut_ptr(), 0, self.instances) }, "read_instance_out")?;
        Ok(out)
    }
    /// `BlockRender::render`: any number of frames in chunks of at most MAX_BLOCK_SIZE
    pub fn render(&mut self, frames: u32) -> Result<Vec<f32>> {
        let mut out = Vec::with_capacity((frames * self.instances * self.channels) as usize);
        let mut pos = 0;
        while pos < frames {
            let chunk = (frames - pos).min(Self::MAX_BLOCK_SIZE);
            self.process_block(chunk)?;
            out.extend(self.out_block()?);
            pos += chunk;
        }
        Ok(out)
    }
    pub fn latency_samples(&self) -> u32 {
        unsafe { sys::oscen_b200_latency_samples(self.raw) }
    }
}
impl Drop for Bank {
    fn drop(&mut self) {
        unsafe { sys::oscen_b200_destroy(self.raw) }
    }
}

/// Hands the body of a `graph!` invocation to the engine as text:
/// ```ignore
/// let graph = b200_graph! {
///     name: Synth;
///     input cutoff: value = 1200.0 [20.0..20000.0, log, ramp: 1323];
///     output audio_out: stream;
///     nodes { osc = PolyBlepOscillator::saw(440.0, 0.5); filter = TptFilter::new(1200.0, 0.707); }
///     connections { osc.output -> filter.input; cutoff -> filter.cutoff; filter.output -> audio_out; }
/// }?;
/// let mut bank = Bank::new(graph, 65_536, 48_000.0, 0)?;
/// ```
/// `stringify!` keeps Rust's token spelling (`2_000.0`, `::`, `->`), which is what the engine's front end lexes.
#[macro_export]
macro_rules! b200_graph {
    ($($body:tt)*) => {
        $crate::Graph::compile(stringify!($($body)*))
    };
}
