/* psxgpu.h — C ABI of the B200-native PSX GPU rasteriser (libpsxgpu.so).
 *
 * Drop-in boundary for DuckStation's software renderer: everything a `GPUBackend`
 * subclass needs in order to replace `GPU_SW` (reference: src/core/gpu_backend.h:30-183,
 * src/core/gpu_sw.h:17-69).  The C++ mirror of that class, `psx::CudaGpuBackend`
 * (duckstation_b200/csrc/backend.h), is a thin layer over these calls; INTEGRATION.md
 * shows the subclass a DuckStation maintainer would add.
 *
 * Conventions: plain pointers and sizes only; every call returns 0 on success or a
 * negative PSXGPU_E_* code and never throws; a context is used by one thread at a time
 * (the reference calls its backend from the video thread only, gpu_backend.h:27-28).
 * There is no CPU fallback: without a CUDA device `psxgpu_create` fails.
 *
 * Command input is the reference's own wire format (include/psxgpu_wire.h ==
 * src/core/video_thread_commands.h): a byte stream of size-prefixed records, exactly what
 * GPUBackend::HandleCommand receives one at a time (src/core/gpu_backend.cpp:383-544).
 */
#ifndef PSXGPU_H
#define PSXGPU_H

#include <stddef.h>
#include <stdint.h>

#include "psxgpu_wire.h"

#ifdef __cplusplus
extern "C" {
#endif

#define PSXGPU_ABI_VERSION 1

enum {
  PSXGPU_OK = 0,
  PSXGPU_E_INVALID = -1,   /* bad argument */
  PSXGPU_E_CUDA = -2,      /* CUDA runtime error; see psxgpu_last_error */
  PSXGPU_E_MALFORMED = -3, /* wire stream record sizes do not add up */
  PSXGPU_E_NOMEM = -4,
  PSXGPU_E_UNSUPPORTED = -5 /* record not allowed on this entry point (e.g. a read-back inside a batch submit) */
};

/* 15-bit scan-out formats (reference picks one from host-device support, gpu_sw.cpp:39-48). */
enum { PSXGPU_FMT_RGBA8 = 0, PSXGPU_FMT_RGB5A1 = 1, PSXGPU_FMT_A1BGR5 = 2, PSXGPU_FMT_RGB565 = 3 };

typedef struct psxgpu_ctx psxgpu_ctx;

/* Replaces the three GPUSettings fields that change GPU_SW output (settings.h:83-85). */
typedef struct psxgpu_opts {
  uint32_t struct_size;      /* sizeof(psxgpu_opts) */
  uint32_t modulation_crop;  /* gpu_modulation_crop */
  uint32_t show_vram;        /* gpu_show_vram: scan-out shows the whole VRAM */
  uint32_t display_format;   /* PSXGPU_FMT_* used for 15-bit scan-out */
  uint32_t clut_slots;       /* CLUT snapshot slots per instance (4..256, 0 = 256) */
  uint32_t encode_threads;   /* host threads for psxgpu_submit_batch (0 = hardware concurrency) */
  uint32_t profile;          /* 1: time every raster wave with CUDA events (psxgpu_get_timing) */
  uint32_t reserved[9];      /* [0] bit 0: no persistent (latency-mode) kernel, bit 1: no instance-mode kernel (one launch per
                                epoch wave; tests and profiling); [1]: where psxgpu_submit_batch encodes —
                                0 auto (on the device for batches of >= 16 instances), 1 host threads, 2 device */
} psxgpu_opts;

typedef struct psxgpu_stats {
  uint64_t wire_commands, packed_prims, epochs, waves;
  uint64_t cuts_raw, cuts_war, cuts_clut;
  uint64_t serial_self_sampling, serial_copy, serial_upload;
  uint64_t clut_ops, clut_cache_hits, rejected;
  uint64_t kernel_launches;     /* our kernels launched since create/reset_stats */
  uint64_t h2d_bytes, d2h_bytes;
  uint64_t flushes;
  uint64_t reserved[7];
} psxgpu_stats;

typedef struct psxgpu_timing {
  float last_flush_device_ms;   /* H2D + setup + all waves of the last flush / run_staged (CUDA events on our stream) */
  float last_kernels_ms;        /* setup + waves only */
  float last_raster_ms;         /* sum over raster waves (profile=1) */
  float last_raster_max_ms;     /* longest single raster launch (profile=1) */
  uint32_t last_raster_launches;
  uint32_t last_waves;
  float last_hash_ms;           /* last psxgpu_hash* call */
  float last_encode_ms;         /* host time spent encoding (host-encoded work) or staging raw streams (device-encoded
                                   batches) during the last submit(s) since the previous flush */
  float last_stage_ms;          /* host time spent packing pinned staging buffers in the last flush */
  float last_scan_ms;           /* batched submit: header scan that sizes the pinned slots */
  float last_pinned_wait_ms;    /* batched submit: host waiting for the device — a pinned staging set still being copied,
                                   or (device-encoded batches) a chunk's framing result / wave table */
  float last_launch_ms;         /* host time building wave descriptors + enqueueing copies and kernels */
  float reserved[4];
} psxgpu_timing;

/* GPUBackend::CreateSoftwareBackend + Initialize(upload_vram=false) (gpu_sw.cpp:34-59, :450-453):
 * n_instances independent 1024x512 VRAMs, cleared, on CUDA device `device`. */
int psxgpu_create(int device, uint32_t n_instances, const psxgpu_opts* opts, psxgpu_ctx** out);
void psxgpu_destroy(psxgpu_ctx* ctx);
const char* psxgpu_last_error(const psxgpu_ctx* ctx);
int psxgpu_abi_version(void);

/* GPUBackend::HandleCommand for a run of records (gpu_backend.cpp:383-544).  Draw / fill / copy /
 * upload / CLUT / drawing-area records are encoded and queued; ClearVRAM, LoadState and ReadVRAM are
 * executed in order (they flush first).  ReadVRAM makes the host shadow current (see psxgpu_shadow_vram);
 * UpdateDisplay scans out into the context's display buffer (psxgpu_display). */
int psxgpu_submit_wire(psxgpu_ctx* ctx, uint32_t instance, const void* wire, size_t bytes);

/* Batched form: streams[i] goes to instance first_instance + i.  Large batches are uploaded raw and encoded on the
 * device (one warp per instance); small ones are encoded on host threads (latency mode).
 * Barrier records (ReadVRAM, LoadState, ClearVRAM, UpdateDisplay) are not allowed here.
 * On PSXGPU_E_MALFORMED / PSXGPU_E_UNSUPPORTED the batch has been applied up to (not including) the chunk that holds the
 * offending stream: earlier chunks are already on the device, the failing chunk is dropped as a whole and its
 * instances' encoder state (drawing area, CLUT snapshot table) is put back to what it was before the chunk. */
int psxgpu_submit_batch(psxgpu_ctx* ctx, uint32_t first_instance, uint32_t count, const void* const* streams,
                        const size_t* bytes);

/* GPUBackend::FlushRender (gpu_backend.h:150): upload queued commands and launch them (asynchronous). */
int psxgpu_flush(psxgpu_ctx* ctx);
/* Wait for everything launched so far. */
int psxgpu_sync(psxgpu_ctx* ctx);

/* psxgpu_submit_batch cuts a batch into chunks of this many instances and launches each chunk as soon as it is
 * encoded, so that host encoding overlaps copies and kernels.  0 = automatic (count/8, at least 128). */
int psxgpu_set_batch_chunk(psxgpu_ctx* ctx, uint32_t instances_per_chunk);

/* Re-run the device part of the most recent flush from the inputs that are still resident in HBM: for a
 * device-encoded batch the raw wire streams (encode + setup + all waves), for host-encoded work the packed command
 * buffers (setup + all waves).  For steady-state measurement; the caller is responsible for the VRAM state it
 * starts from. */
int psxgpu_run_staged(psxgpu_ctx* ctx);

/* GPUBackend::ReadVRAM contract (gpu_backend.h:152, gpu_hw.cpp:3472-3515): after the call the pixels of the
 * rectangle (wrapping like GPUREAD, gpu.cpp:1896-1929) are in dst, row pitch dst_pitch_px pixels. */
int psxgpu_read_vram(psxgpu_ctx* ctx, uint32_t instance, uint32_t x, uint32_t y, uint32_t w, uint32_t h, uint16_t* dst,
                     uint32_t dst_pitch_px);
/* The reference exposes VRAM as the global g_vram; this is its stand-in: a 1024x512 host copy refreshed by
 * ReadVRAM records / psxgpu_read_vram(…, shadow) for `instance`. */
uint16_t* psxgpu_shadow_vram(psxgpu_ctx* ctx, uint32_t instance);

/* GPUBackend::ClearVRAM (gpu_sw.cpp:61-65) and LoadState (gpu_sw.cpp:67-71). */
int psxgpu_clear_vram(psxgpu_ctx* ctx, uint32_t instance);
int psxgpu_load_state(psxgpu_ctx* ctx, uint32_t instance, const uint16_t* vram_1024x512, const uint16_t* clut_256);
/* DoMemoryState (gpu_sw.cpp:73-84): VRAM + the current 256-entry CLUT cache. */
int psxgpu_save_state(psxgpu_ctx* ctx, uint32_t instance, uint16_t* vram_1024x512, uint16_t* clut_256);

/* GPU_SW::UpdateDisplay minus presentation (gpu_sw.cpp:391-448).  Writes width*height pixels (RGBA8 for 24-bit
 * sources, opts.display_format otherwise) to dst with row pitch dst_stride bytes.  Returns 1 if the display is
 * disabled (nothing written). */
int psxgpu_scanout(psxgpu_ctx* ctx, uint32_t instance, const psx_update_display* cmd, void* dst, uint32_t dst_stride,
                   uint32_t* out_width, uint32_t* out_height);
/* Result of the last UpdateDisplay record seen by psxgpu_submit_wire. */
const void* psxgpu_display(psxgpu_ctx* ctx, uint32_t* width, uint32_t* height, uint32_t* stride, uint32_t* is_rgba8);

/* ---- presenter ------------------------------------------------------------------------------
 * GPU_SW::UpdateDisplay including the steps it hands to VideoPresenter (gpu_sw.cpp:386-441): scan-out, then
 * ApplyChromaSmoothing on 24-bit frames when enabled (video_presenter.cpp:1385-1413), then Deinterlace(field) on
 * interlaced display (video_presenter.cpp:1208-1371).  The reference runs these as fragment shaders
 * (video_shadergen.cpp:166-330); here they are CUDA kernels on RGBA8 frames that never leave the device in between.
 * The object owns what the reference keeps in VideoPresenter's statics: the deinterlace target, up to four field
 * buffers and the current-buffer index.  Parity status: pinned by our reading of the shader text. */
typedef struct psxgpu_presenter psxgpu_presenter;
enum { /* DisplayDeinterlacingMode, src/core/types.h:81-89 */
  PSXGPU_DEINT_DISABLED = 0, PSXGPU_DEINT_WEAVE = 1, PSXGPU_DEINT_BLEND = 2, PSXGPU_DEINT_ADAPTIVE = 3,
  PSXGPU_DEINT_PROGRESSIVE = 4
};
int psxgpu_presenter_create(psxgpu_ctx* ctx, uint32_t deinterlace_mode, uint32_t chroma_smoothing, psxgpu_presenter** out);
void psxgpu_presenter_destroy(psxgpu_presenter* p);
/* One UpdateDisplay: the frame that would be shown, tightly packed RGBA8, into dst (host).  Returns 0, 1 if nothing
 * is shown (display disabled and no deinterlaced frame to keep), or a negative PSXGPU_E_* code.  The frame is always
 * scanned out as RGBA8, whatever psxgpu_opts.display_format says. */
int psxgpu_present(psxgpu_presenter* p, uint32_t instance, const psx_update_display* cmd, void* dst, size_t dst_bytes,
                   uint32_t* out_width, uint32_t* out_height);

/* vramhash64 of instances [first, first+count): sum over 32-bit words i of mix64((i<<32)|word) (include
 * psxgpu_hash.h semantics in DESIGN.md §7).  Flushes and waits. */
int psxgpu_hash(psxgpu_ctx* ctx, uint32_t first_instance, uint32_t count, uint64_t* out_host);
/* Same, result left in device memory (e.g. a torch tensor) for a following NCCL gather; waits for completion. */
int psxgpu_hash_device(psxgpu_ctx* ctx, uint32_t first_instance, uint32_t count, void* out_device);

int psxgpu_get_stats(psxgpu_ctx* ctx, psxgpu_stats* out);
int psxgpu_reset_stats(psxgpu_ctx* ctx);
int psxgpu_get_timing(psxgpu_ctx* ctx, psxgpu_timing* out);
/* Durations (ms) of the raster launches of the last profiled flush / run_staged; returns the count. */
int psxgpu_get_raster_launch_ms(psxgpu_ctx* ctx, float* out, uint32_t max_count);

/* ---- GP0/GP1 word-level front-end (host only, no device work) -------------------------------
 * Restates the reference's port decoder, src/core/gpu.cpp: WriteGP0/ExecuteCommands (:2618-2812), the draw-state
 * registers (:2866-2954), the primitive and transfer handlers (:3099-3891), WriteGP1 (:1931-2106) and the display
 * half of UpdateCRTCDisplayParameters / UpdateDisplay (:1243-1431, :2451-2481).  Output is the backend record
 * stream psxgpu_submit_wire consumes.  Parity status: unpinned. */
typedef struct psxgpu_frontend psxgpu_frontend;
typedef struct psxgpu_frontend_stats {
  uint64_t gp0_words, gp1_words, commands, unknown_commands, culled_primitives, frames;
  uint64_t idle; /* 1: FIFO empty and no transfer / polyline in flight */
  uint64_t reserved[5];
} psxgpu_frontend_stats;
/* flags: PSXGPU_FE_*; crop_mode: 0 none, 1 overscan (the reference's default), 2 borders (settings.h DisplayCropMode). */
enum {
  PSXGPU_FE_PAL = 1,       /* PAL console region (System::IsPALRegion) */
  PSXGPU_FE_INTERLACED = 2, /* display_deinterlacing_mode != Progressive: interlaced rendering + field scan-out on */
  PSXGPU_FE_FIELD_PER_VSYNC = 4 /* flip the interlace field and the active-line bit at every psxgpu_frontend_vsync: a stand-in
                                   for the beam-accurate CRTC of a running console (gpu.cpp:1679-1721).  Without it the two
                                   bits only change through GP1 / a loaded state — what the reference itself does while it
                                   replays a .psxgpu trace (GPU::ProcessGPUDumpPacket, gpu.cpp:4151-4168: the CRTC event is
                                   not running), and what the decoder parity tests pin. */
};
int psxgpu_frontend_create(uint32_t flags, uint32_t crop_mode, psxgpu_frontend** out);
void psxgpu_frontend_destroy(psxgpu_frontend* fe);
int psxgpu_frontend_gp0(psxgpu_frontend* fe, const uint32_t* words, size_t count); /* GPU::WriteRegister(0) / DMA */
int psxgpu_frontend_gp1(psxgpu_frontend* fe, uint32_t value);                     /* GPU::WriteRegister(4) */
int psxgpu_frontend_vsync(psxgpu_frontend* fe);        /* end of frame: emits UpdateDisplay, flips the field */
int psxgpu_frontend_finish_read(psxgpu_frontend* fe);  /* the CPU has drained GPUREAD after a VRAM->CPU command */
/* Records decoded so far (valid until the next frontend call); the caller submits them and clears. */
const void* psxgpu_frontend_records(psxgpu_frontend* fe, size_t* bytes);
int psxgpu_frontend_clear(psxgpu_frontend* fe);
int psxgpu_frontend_get_stats(psxgpu_frontend* fe, psxgpu_frontend_stats* out);
/* Feeds a whole `.psxgpu` trace (see psxgpu_replay_dump) through the front-end; the records of all frames, each
 * closed by its UpdateDisplay, accumulate in psxgpu_frontend_records.  Returns frames or PSXGPU_E_MALFORMED. */
long psxgpu_frontend_feed_dump(psxgpu_frontend* fe, const void* data, size_t bytes);

/* Replays a `.psxgpu` trace (reference: src/core/gpu_dump.cpp, GPUDump::Player) into `instance`: every frame's
 * records go through psxgpu_submit_wire, then on_frame(user, frame_index) is called (the display buffer of that
 * frame is current, psxgpu_display).  Returns the number of frames or a negative PSXGPU_E_* code. */
typedef void (*psxgpu_frame_callback)(void* user, long frame);
long psxgpu_replay_dump(psxgpu_ctx* ctx, uint32_t instance, const void* data, size_t bytes, uint32_t flags,
                        uint32_t crop_mode, psxgpu_frame_callback on_frame, void* user);

/* ---- save states ---------------------------------------------------------------------------
 * The GPU's part of a DuckStation save state: what GPU::DoState reads / writes (src/core/gpu.cpp:574-726) between the
 * "GPU" and "CDROM" markers of System::DoState (src/core/system.cpp:2497-2500), in StateWrapper encoding
 * (src/util/state_wrapper.h).  Loading sets the front-end's registers / FIFO / transfer state and pushes VRAM + the
 * CLUT cache into `instance` like the GPUBackendLoadStateCommand the reference issues; saving writes a version-86
 * section (timing fields zero, empty texture-cache tail).  Parity status: unpinned (no save state to read here;
 * DESIGN.md §13). */
typedef struct psxgpu_state_file { /* the fields of SAVE_STATE_HEADER (src/core/save_state_version.h:14-56) a loader needs */
  uint32_t version;
  uint32_t data_compression; /* 0 none, 1 deflate, 2 zstd, 3 xz: decompress data[offset_to_data ..] before searching */
  uint32_t data_compressed_size, data_uncompressed_size, offset_to_data;
  char title[128];
  char serial[32];
} psxgpu_state_file;
int psxgpu_state_file_info(const void* file, size_t bytes, psxgpu_state_file* out);
/* Offset of the GPU section inside the (decompressed) state data: the byte after the "GPU" marker whose section
 * parses completely and is followed by the "CDROM" marker (or by the end of the buffer).  Negative if none. */
long psxgpu_state_find_gpu_section(const void* state, size_t bytes, uint32_t version);
/* ctx may be NULL (registers only).  consumed: bytes of the section. */
int psxgpu_frontend_load_state(psxgpu_frontend* fe, psxgpu_ctx* ctx, uint32_t instance, const void* section, size_t bytes,
                               uint32_t version, size_t* consumed);
/* dst NULL: only *written is set (size query).  ctx NULL: registers only, VRAM and the CLUT cache are written as zero. */
int psxgpu_frontend_save_state(psxgpu_frontend* fe, psxgpu_ctx* ctx, uint32_t instance, void* dst, size_t capacity,
                               size_t* written);

#ifdef __cplusplus
}
#endif
#endif /* PSXGPU_H */
