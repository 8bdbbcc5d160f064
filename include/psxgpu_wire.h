/* psxgpu_wire.h — the backend command wire format, as plain C.
 *
 * These are layout-compatible restatements of the POD command structs the
 * reference's GP0 decoder pushes to a GPUBackend
 * (reference: src/core/video_thread_commands.h:27-56 command types, :58-69 header,
 * :131-362 payloads; register types src/core/gpu_types.h:108-266).  A stream of
 * these records, each `size` bytes long (size is a multiple of 16, header
 * included, video_thread_commands.h:63-68), is what `psxgpu_submit_wire()` and
 * the C++ `CudaGpuBackend::HandleCommand()` consume, so a ring slot of the
 * reference can be handed over without conversion.
 *
 * Offsets were probed from the reference headers with g++ 13 (x86-64 SysV) and
 * are pinned again by static_asserts in oracle/ref_build/ref_harness.cpp, which
 * includes both this file and the reference header.
 *
 * Bitfields of the reference are expressed as flag bytes + masks because the
 * in-memory order of C bitfields is what the reference ABI fixes, not the
 * declaration.
 */
#ifndef PSXGPU_WIRE_H
#define PSXGPU_WIRE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
  PSX_VRAM_WIDTH = 1024,
  PSX_VRAM_HEIGHT = 512,
  PSX_VRAM_PIXELS = PSX_VRAM_WIDTH * PSX_VRAM_HEIGHT,
  PSX_VRAM_BYTES = PSX_VRAM_PIXELS * 2,
  PSX_CLUT_ENTRIES = 256,
  PSX_MAX_PRIM_WIDTH = 1024,
  PSX_MAX_PRIM_HEIGHT = 512
};

/* VideoThreadCommandType values (video_thread_commands.h:27-56).  Only the
 * ones that reach GPUBackend::HandleCommand (gpu_backend.cpp:383-544). */
enum psx_cmd_type {
  PSX_CMD_CLEAR_VRAM = 7,
  PSX_CMD_CLEAR_DISPLAY = 8,
  PSX_CMD_UPDATE_DISPLAY = 9,
  PSX_CMD_SUBMIT_FRAME = 10,
  PSX_CMD_BUFFER_SWAPPED = 11,
  PSX_CMD_LOAD_STATE = 12,
  PSX_CMD_READ_VRAM = 15,
  PSX_CMD_FILL_VRAM = 16,
  PSX_CMD_UPDATE_VRAM = 17,
  PSX_CMD_COPY_VRAM = 18,
  PSX_CMD_SET_DRAWING_AREA = 19,
  PSX_CMD_UPDATE_CLUT = 20,
  PSX_CMD_CLEAR_CACHE = 21,
  PSX_CMD_DRAW_POLYGON = 22,
  PSX_CMD_DRAW_PRECISE_POLYGON = 23,
  PSX_CMD_DRAW_RECTANGLE = 24,
  PSX_CMD_DRAW_LINE = 25,
  PSX_CMD_DRAW_PRECISE_LINE = 26
};

/* VideoThreadCommand (video_thread_commands.h:58-69): 8 bytes. */
typedef struct psx_cmd_header {
  uint32_t size; /* whole record incl. header, multiple of 16 */
  uint8_t type;  /* enum psx_cmd_type */
  uint8_t pad_[3];
} psx_cmd_header;

/* GPUBackendDrawCommand flag bits (video_thread_commands.h:251-266). */
enum {
  /* flags0, byte 8 */
  PSX_DF_INTERLACED = 0x01,
  PSX_DF_ACTIVE_LINE_LSB = 0x02,
  PSX_DF_SET_MASK = 0x04,
  PSX_DF_CHECK_MASK = 0x08,
  PSX_DF_TEXTURE = 0x10,
  PSX_DF_RAW_TEXTURE = 0x20,
  PSX_DF_TRANSPARENCY = 0x40,
  PSX_DF_SHADING = 0x80,
  /* flags1, byte 9 */
  PSX_DF1_QUAD = 0x01,
  PSX_DF1_DITHER = 0x02,
  PSX_DF1_VALID_W = 0x04
};

/* GPUDrawModeReg fields (gpu_types.h:208-237). */
#define PSX_DM_PAGE_X(bits) ((uint32_t)((bits) & 0xFu))       /* *64  */
#define PSX_DM_PAGE_Y(bits) ((uint32_t)(((bits) >> 4) & 1u))  /* *256 */
#define PSX_DM_TRANSPARENCY(bits) ((uint32_t)(((bits) >> 5) & 3u))
#define PSX_DM_TEXMODE(bits) ((uint32_t)(((bits) >> 7) & 3u))
#define PSX_DM_DITHER(bits) ((uint32_t)(((bits) >> 9) & 1u))

/* GPUTexturePaletteReg (gpu_types.h:239-250): x = bits 0-5 (*16), y = bits 6-14. */
#define PSX_PAL_X(bits) (((uint32_t)(bits) & 0x3Fu) * 16u)
#define PSX_PAL_Y(bits) (((uint32_t)(bits) >> 6) & 0x1FFu)

/* GPUTextureWindow (gpu_types.h:252-266). */
typedef struct psx_texture_window {
  uint8_t and_x, and_y, or_x, or_y;
} psx_texture_window;

/* GPUBackendDrawCommand (video_thread_commands.h:249-276): 20 bytes. */
typedef struct psx_draw_header {
  psx_cmd_header h;
  uint8_t flags0;
  uint8_t flags1;
  uint16_t num_vertices;
  uint16_t draw_mode; /* GPUDrawModeReg.bits */
  uint16_t palette;   /* GPUTexturePaletteReg.bits — ignored by the SW renderer */
  psx_texture_window window;
} psx_draw_header;

/* GPUBackendDrawPolygonCommand::Vertex (video_thread_commands.h:280-299): 16 bytes. */
typedef struct psx_poly_vertex {
  int32_t x, y;
  uint8_t r, g, b, a;
  uint8_t u, v;
  uint8_t pad_[2];
} psx_poly_vertex;

typedef struct psx_draw_polygon {
  psx_draw_header d;
  psx_poly_vertex vertices[4]; /* num_vertices (3 or 4) are valid */
} psx_draw_polygon;

/* GPUBackendDrawPrecisePolygonCommand (video_thread_commands.h:304-317). */
typedef struct psx_precise_poly_vertex {
  float x, y, w;
  int32_t native_x, native_y;
  uint32_t color;
  uint16_t texcoord;
  uint8_t pad_[2];
} psx_precise_poly_vertex;

typedef struct psx_draw_precise_polygon {
  psx_draw_header d;
  psx_draw_header params; /* unused by the SW renderer (gpu_sw.cpp:118-139) */
  psx_precise_poly_vertex vertices[4];
} psx_draw_precise_polygon;

/* GPUBackendDrawRectangleCommand (video_thread_commands.h:319-325): 40 bytes. */
typedef struct psx_draw_rectangle {
  psx_draw_header d;
  uint16_t width, height;
  uint16_t texcoord; /* u | v<<8 */
  uint8_t pad_[2];
  int32_t x, y;
  uint32_t color;
} psx_draw_rectangle;

/* GPUBackendDrawLineCommand::Vertex (video_thread_commands.h:329-347): 12 bytes.
 * Lines arrive as vertex PAIRS (polylines already expanded, gpu.cpp:3644-3691). */
typedef struct psx_line_vertex {
  int32_t x, y;
  uint8_t r, g, b, a;
} psx_line_vertex;

typedef struct psx_draw_line {
  psx_draw_header d;
  psx_line_vertex vertices[2]; /* num_vertices follow, flexible */
} psx_draw_line;

typedef struct psx_precise_line_vertex {
  float x, y, w;
  int32_t native_x, native_y;
  uint32_t color;
} psx_precise_line_vertex;

/* GPUBackendFillVRAMCommand (video_thread_commands.h:204-213): 24 bytes. */
typedef struct psx_fill_vram {
  psx_cmd_header h;
  uint16_t x, y, width, height;
  uint32_t color; /* RGBA8888; bit 24 becomes bit 15 (gpu_helpers.h:32-39) */
  uint8_t interlaced_rendering;
  uint8_t active_line_lsb;
  uint8_t pad_[2];
} psx_fill_vram;

/* GPUBackendUpdateVRAMCommand (video_thread_commands.h:215-224): 18-byte head,
 * width*height u16 pixels follow at offset 18. */
typedef struct psx_update_vram {
  psx_cmd_header h;
  uint16_t x, y, width, height;
  uint8_t set_mask_while_drawing;
  uint8_t check_mask_before_draw;
  uint16_t data[1]; /* flexible: width*height */
} psx_update_vram;
#define PSX_UPDATE_VRAM_DATA_OFFSET 18u

/* GPUBackendCopyVRAMCommand (video_thread_commands.h:226-236): 24 bytes. */
typedef struct psx_copy_vram {
  psx_cmd_header h;
  uint16_t src_x, src_y, dst_x, dst_y, width, height;
  uint8_t set_mask_while_drawing;
  uint8_t check_mask_before_draw;
  uint8_t pad_[2];
} psx_copy_vram;

/* GPUDrawingArea (gpu_types.h:107-114): inclusive bounds. */
typedef struct psx_drawing_area {
  uint32_t left, top, right, bottom;
} psx_drawing_area;

/* GPUBackendSetDrawingAreaCommand (video_thread_commands.h:238-241): 24 bytes. */
typedef struct psx_set_drawing_area {
  psx_cmd_header h;
  psx_drawing_area new_area;
} psx_set_drawing_area;

/* GPUBackendUpdateCLUTCommand (video_thread_commands.h:243-247): 12 bytes. */
typedef struct psx_update_clut {
  psx_cmd_header h;
  uint16_t reg; /* GPUTexturePaletteReg.bits */
  uint8_t clut_is_8bit;
  uint8_t pad_;
} psx_update_clut;

/* GPUBackendReadVRAMCommand (video_thread_commands.h:196-202): 16 bytes. */
typedef struct psx_read_vram {
  psx_cmd_header h;
  uint16_t x, y, width, height;
} psx_read_vram;

/* GPUBackendUpdateDisplayCommand flag bits, byte 31 (video_thread_commands.h:178-185). */
enum {
  PSX_UD_INTERLACED_ENABLED = 0x01,
  PSX_UD_INTERLACED_FIELD = 0x02,
  PSX_UD_INTERLACED_INTERLEAVED = 0x04,
  PSX_UD_480I_MODE = 0x08,
  PSX_UD_DISPLAY_24BIT = 0x10,
  PSX_UD_DISPLAY_DISABLED = 0x20,
  PSX_UD_SUBMIT_FRAME = 0x40
};

/* GPUBackendUpdateDisplayCommand (video_thread_commands.h:162-188): 64 bytes. */
typedef struct psx_update_display {
  psx_cmd_header h;
  uint16_t display_width, display_height;
  uint16_t display_origin_left, display_origin_top;
  uint16_t display_vram_left, display_vram_top;
  uint16_t display_vram_width, display_vram_height;
  float display_pixel_aspect_ratio;
  uint16_t X;
  uint8_t gpu_busy_pct;
  uint8_t flags;
  uint8_t frame_[32]; /* GPUBackendFramePresentationParameters — presentation only */
} psx_update_display;

/* GPUBackendLoadStateCommand (video_thread_commands.h:131-138). */
typedef struct psx_load_state {
  psx_cmd_header h;
  uint16_t vram_data[PSX_VRAM_PIXELS];
  uint16_t clut_data[PSX_CLUT_ENTRIES];
  uint32_t texture_cache_state_version;
  uint32_t texture_cache_state_size;
} psx_load_state;

#ifdef __cplusplus
} /* extern "C" */

static_assert(sizeof(psx_cmd_header) == 8, "header");
static_assert(sizeof(psx_draw_header) == 20, "draw header");
static_assert(offsetof(psx_draw_header, flags0) == 8 && offsetof(psx_draw_header, num_vertices) == 10 &&
                offsetof(psx_draw_header, draw_mode) == 12 && offsetof(psx_draw_header, window) == 16,
              "draw header");
static_assert(sizeof(psx_poly_vertex) == 16 && offsetof(psx_poly_vertex, u) == 12, "poly vertex");
static_assert(sizeof(psx_precise_poly_vertex) == 28 && offsetof(psx_precise_poly_vertex, native_x) == 12 &&
                offsetof(psx_precise_poly_vertex, texcoord) == 24,
              "precise vertex");
static_assert(offsetof(psx_draw_precise_polygon, vertices) == 40, "precise polygon");
static_assert(sizeof(psx_draw_rectangle) == 40 && offsetof(psx_draw_rectangle, x) == 28 &&
                offsetof(psx_draw_rectangle, color) == 36,
              "rect");
static_assert(sizeof(psx_line_vertex) == 12 && sizeof(psx_precise_line_vertex) == 24, "line vertex");
static_assert(sizeof(psx_fill_vram) == 24 && offsetof(psx_fill_vram, color) == 16 &&
                offsetof(psx_fill_vram, active_line_lsb) == 21,
              "fill");
static_assert(offsetof(psx_update_vram, data) == PSX_UPDATE_VRAM_DATA_OFFSET, "update");
static_assert(sizeof(psx_copy_vram) == 24 && offsetof(psx_copy_vram, check_mask_before_draw) == 21, "copy");
static_assert(sizeof(psx_set_drawing_area) == 24, "area");
static_assert(sizeof(psx_update_clut) == 12 && offsetof(psx_update_clut, clut_is_8bit) == 10, "clut");
static_assert(sizeof(psx_read_vram) == 16, "read");
static_assert(sizeof(psx_update_display) == 64 && offsetof(psx_update_display, X) == 28 &&
                offsetof(psx_update_display, flags) == 31,
              "display");
static_assert(offsetof(psx_load_state, clut_data) == 1048584 && sizeof(psx_load_state) == 1049104, "load state");
#endif

#endif /* PSXGPU_WIRE_H */
