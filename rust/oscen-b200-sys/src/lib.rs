//! Raw bindings to `include/oscen_b200.h` (ABI version 1).  One item per C entry point, in header order.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

pub const OSCEN_B200_ABI_VERSION: c_int = 1;
pub const OSCEN_B200_MAX_BLOCK: u32 = 512;
pub const OSCEN_B200_EVENT_GATE: u32 = 0xFFFF_FFFF;
pub const OSCEN_B200_GRAPH_CUSTOM: u32 = 20;
pub const OSCEN_B200_PLAN_FUSED: u32 = 0;
pub const OSCEN_B200_PLAN_NODES: u32 = 1;
pub const OSCEN_B200_KEEP_VOICE_OUT: u32 = 1;
pub const OSCEN_B200_KEEP_INSTANCE_OUT: u32 = 2;
pub const OSCEN_B200_E_JIT: c_int = -7;

#[repr(C)]
pub struct oscen_b200_engine {
    _private: [u8; 0],
}
#[repr(C)]
pub struct oscen_b200_graph {
    _private: [u8; 0],
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct oscen_b200_config {
    pub graph_kind: u32,
    pub n_voices: u32,
    pub voices_per_instance: u32,
    pub sample_rate: f32,
    pub device: i32,
    pub plan: u32,
    pub keep: u32,
    pub reserved: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct oscen_b200_event {
    pub voice: u32,
    pub frame_offset: u32,
    pub input_id: u32,
    pub scalar: f32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct oscen_b200_midi_msg {
    pub instance: u32,
    pub frame_offset: u32,
    pub bytes: [u8; 3],
    pub len: u8,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct oscen_b200_input_info {
    pub name: [c_char; 64],
    pub default_value: f32,
    pub has_range: u32,
    pub range_min: f32,
    pub range_max: f32,
    pub logarithmic: u32,
    pub ramp_frames: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct oscen_b200_kind_tables {
    pub n_inputs: u32,
    pub n_state: u32,
    pub channels: u32,
    pub dyn_smem_bytes: u32,
    pub has_stream_in: u32,
    pub reserved: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct oscen_b200_kernel_info {
    pub source_len: u32,
    pub n_state: u32,
    pub dyn_smem_bytes: u32,
    pub inner_factor: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct oscen_b200_stats {
    pub kernel_launches: u64,
    pub graph_launches: u64,
    pub h2d_bytes: u64,
    pub d2h_bytes: u64,
    pub blocks: u64,
    pub events_dropped: u64,
}

extern "C" {
    pub fn oscen_b200_abi_version() -> c_int;
    pub fn oscen_b200_device_count() -> c_int;
    pub fn oscen_b200_create(cfg: *const oscen_b200_config, out: *mut *mut oscen_b200_engine) -> c_int;
    pub fn oscen_b200_destroy(e: *mut oscen_b200_engine);
    pub fn oscen_b200_last_error(e: *const oscen_b200_engine) -> *const c_char;
    pub fn oscen_b200_set_value(e: *mut oscen_b200_engine, input_id: u32, first_voice: u32, count: u32, v: *const f32, broadcast: c_int) -> c_int;
    pub fn oscen_b200_set_value_ramp(e: *mut oscen_b200_engine, input_id: u32, target: f32, frames: u32) -> c_int;
    pub fn oscen_b200_push_events(e: *mut oscen_b200_engine, ev: *const oscen_b200_event, n: usize) -> c_int;
    pub fn oscen_b200_push_midi(e: *mut oscen_b200_engine, instance: u32, frame_offset: u32, bytes: *const u8, len: u32) -> c_int;
    pub fn oscen_b200_push_midi_batch(e: *mut oscen_b200_engine, msgs: *const oscen_b200_midi_msg, n: usize) -> c_int;
    pub fn oscen_b200_stream_in(e: *mut oscen_b200_engine, input_id: u32, src: *const f32, frames: u32, is_device: c_int) -> c_int;
    pub fn oscen_b200_stream_in_lcg(e: *mut oscen_b200_engine, input_id: u32, seed: u64, voice_offset: u32) -> c_int;
    pub fn oscen_b200_process_block(e: *mut oscen_b200_engine, frames: u32) -> c_int;
    pub fn oscen_b200_synchronize(e: *mut oscen_b200_engine) -> c_int;
    pub fn oscen_b200_read_voice_out(e: *mut oscen_b200_engine, dst: *mut f32, first_voice: u32, count: u32) -> c_int;
    pub fn oscen_b200_read_instance_out(e: *mut oscen_b200_engine, dst: *mut f32, first_instance: u32, count: u32) -> c_int;
    pub fn oscen_b200_read_mix(e: *mut oscen_b200_engine, dst: *mut f32) -> c_int;
    pub fn oscen_b200_channels(e: *const oscen_b200_engine) -> u32;
    pub fn oscen_b200_latency_samples(e: *const oscen_b200_engine) -> u32;
    pub fn oscen_b200_n_inputs(graph_kind: u32) -> u32;
    pub fn oscen_b200_state_bytes_per_voice(e: *const oscen_b200_engine) -> usize;
    pub fn oscen_b200_state_download(e: *mut oscen_b200_engine, dst: *mut c_void, cap: usize) -> c_int;
    pub fn oscen_b200_plan_describe(graph_kind: u32, dst: *mut c_char, cap: usize) -> c_int;
    pub fn oscen_b200_graph_compile(dsl_text: *const c_char, out: *mut *mut oscen_b200_graph, err: *mut c_char, err_cap: usize) -> c_int;
    pub fn oscen_b200_graph_compile_named(dsl_text: *const c_char, graph_name: *const c_char, out: *mut *mut oscen_b200_graph, err: *mut c_char, err_cap: usize) -> c_int;
    pub fn oscen_b200_graph_free(g: *mut oscen_b200_graph);
    pub fn oscen_b200_graph_describe(g: *const oscen_b200_graph, dst: *mut c_char, cap: usize) -> c_int;
    pub fn oscen_b200_graph_n_inputs(g: *const oscen_b200_graph) -> u32;
    pub fn oscen_b200_graph_voices_per_instance(g: *const oscen_b200_graph) -> u32;
    pub fn oscen_b200_graph_channels(g: *const oscen_b200_graph) -> u32;
    pub fn oscen_b200_graph_has_midi(g: *const oscen_b200_graph) -> c_int;
    pub fn oscen_b200_graph_input_id(g: *const oscen_b200_graph, name: *const c_char) -> c_int;
    pub fn oscen_b200_graph_input_info(g: *const oscen_b200_graph, input_id: u32, out: *mut oscen_b200_input_info) -> c_int;
    pub fn oscen_b200_create_from_graph(g: *const oscen_b200_graph, cfg: *const oscen_b200_config, out: *mut *mut oscen_b200_engine) -> c_int;
    pub fn oscen_b200_specialize(e: *mut oscen_b200_engine, uniform_mask: *mut u32) -> c_int;
    pub fn oscen_b200_graph_codegen(g: *const oscen_b200_graph, dst: *mut c_char, cap: usize) -> c_int;
    pub fn oscen_b200_graph_jit_check(g: *const oscen_b200_graph, log: *mut c_char, log_cap: usize) -> c_int;
    pub fn oscen_b200_kind_info(graph_kind: u32, sample_rate: f32, info: *mut oscen_b200_kind_tables, defaults: *mut f32, state_words: *mut u32, state_cap: usize) -> c_int;
    pub fn oscen_b200_graph_n_assets(g: *const oscen_b200_graph) -> u32;
    pub fn oscen_b200_graph_asset_id(g: *const oscen_b200_graph, name: *const c_char) -> c_int;
    pub fn oscen_b200_load_asset(e: *mut oscen_b200_engine, asset_id: u32, interleaved: *const f32, frames: u32, channels: u32) -> c_int;
    pub fn oscen_b200_graph_kernel_info(g: *const oscen_b200_graph, sample_rate: f32, info: *mut oscen_b200_kernel_info, source: *mut c_char, source_cap: usize, state_words: *mut u32, state_cap: usize) -> c_int;
    pub fn oscen_b200_graph_kernel_source(g: *const oscen_b200_graph, sample_rate: f32, uniform_mask: u32, dst: *mut c_char, cap: usize) -> c_int;
    pub fn oscen_b200_get_stats(e: *const oscen_b200_engine, out: *mut oscen_b200_stats) -> c_int;
    pub fn oscen_b200_last_block_ms(e: *mut oscen_b200_engine, ms: *mut f32) -> c_int;
    pub fn oscen_b200_enable_timing(e: *mut oscen_b200_engine, on: c_int) -> c_int;
    pub fn oscen_b200_last_kernel_ms(e: *mut oscen_b200_engine, ms: *mut f32) -> c_int;
    pub fn oscen_b200_timing_drain(e: *mut oscen_b200_engine, block_ms: *mut f32, kernel_ms: *mut f32, cap: u32) -> c_int;
    pub fn oscen_b200_last_kernel_name(e: *const oscen_b200_engine) -> *const c_char;
    pub fn oscen_b200_l2_flush(e: *mut oscen_b200_engine) -> c_int;
    pub fn oscen_b200_comm_unique_id(id128: *mut c_void) -> c_int;
    pub fn oscen_b200_comm_init(e: *mut oscen_b200_engine, rank: c_int, n_ranks: c_int, id128: *const c_void) -> c_int;
    pub fn oscen_b200_comm_p2p_export(e: *mut oscen_b200_engine, n_ranks: c_int, handle64: *mut c_void) -> c_int;
    pub fn oscen_b200_comm_p2p_import(e: *mut oscen_b200_engine, rank: c_int, n_ranks: c_int, handles: *const c_void) -> c_int;
    pub fn oscen_b200_comm_p2p_close(e: *mut oscen_b200_engine) -> c_int;
}
