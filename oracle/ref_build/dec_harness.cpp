// dec_harness.cpp — TEST INFRASTRUCTURE.  Link stubs + C entry points that let the reference's own GP0/GP1 decoder
// (src/core/gpu.cpp, compiled IN PLACE from /root/reference together with util/state_wrapper.cpp,
// common/small_string.cpp and dep/fmt/src/format.cc — nothing copied) run without the emulator around it:
//
//   * GPU::ProcessGPUDumpPacket (gpu.cpp:4087, the reference's own .psxgpu replay entry) is fed GP0 / GP1 / VSync
//     packets; the backend records the decoder emits (GPUBackend::New*Command + PushCommand) land in a capture arena
//     instead of the video thread's ring and are handed back byte for byte — they are what include/psxgpu_wire.h
//     restates and what csrc/gp0_frontend.cpp must produce from the same words;
//   * GPU::DoState (gpu.cpp:574) writes / reads the "GPU" save-state section through the real StateWrapper
//     (row (f)4).
//
// Everything below is a STUB of a subsystem the decoder only notifies (timing events, DMA, IRQs, timers, host, ImGui,
// texture cache, dump recorder): no-ops or constants, listed one by one so that the stub surface is on record
// (oracle/ref_build/README_decoder.md).  Never linked into the product.
#include "core/bus.h"
#include "core/cpu_pgxp.h"
#include "core/dma.h"
#include "core/gpu.h"
#include "core/gpu_backend.h"
#include "core/gpu_dump.h"
#include "core/gpu_hw_texture_cache.h"
#include "core/host.h"
#include "core/interrupt_controller.h"
#include "core/performance_counters.h"
#include "core/settings.h"
#include "core/system.h"
#include "core/system_private.h"
#include "core/timers.h"
#include "core/timing_event.h"
#include "core/video_thread.h"
#include "core/video_thread_commands.h"

#include "util/image.h"
#include "util/state_wrapper.h"

#include "core/core.h"

#include "util/imgui_manager.h"
#include "util/translation.h"
#include "util/window_info.h"

#include "common/error.h"
#include "common/file_system.h"
#include "common/log.h"
#include "common/path.h"
#include "common/string_util.h"

#include "imgui.h"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

// ---------------------------------------------------------------------------------------------- capture arena
namespace {
std::vector<u8> s_arena;        // finished records, back to back (each VideoThreadCommand::size bytes, 16-byte multiples)
std::vector<u8> s_scratch;      // the record being built (New*Command ... PushCommand)
bool s_replaying = true;
u32 s_frame_number = 0;
GlobalTicks s_global_ticks = 0;
bool s_pal = false;
} // namespace

VideoThreadCommand* VideoThread::AllocateCommand(VideoThreadCommandType command, u32 size)
{
  size = VideoThreadCommand::AlignCommandSize(size);
  s_scratch.assign(size, 0);
  VideoThreadCommand* cmd = reinterpret_cast<VideoThreadCommand*>(s_scratch.data());
  cmd->type = command;
  cmd->size = size;
  return cmd;
}
void VideoThread::PushCommand(VideoThreadCommand* cmd)
{
  const u8* p = reinterpret_cast<const u8*>(cmd);
  s_arena.insert(s_arena.end(), p, p + cmd->size);
}
void VideoThread::PushCommandAndWakeThread(VideoThreadCommand* cmd)
{
  PushCommand(cmd);
}
void VideoThread::SyncThread(bool)
{
}
const WindowInfo& VideoThread::GetRenderWindowInfo()
{
  static WindowInfo wi;
  return wi;
}

// GPUBackend command constructors: sizes as src/core/gpu_backend.cpp:130-246
#define NEW_CMD(T, TYPE, SIZE) static_cast<T*>(VideoThread::AllocateCommand(VideoThreadCommandType::TYPE, SIZE))
VideoThreadCommand* GPUBackend::NewClearVRAMCommand() { return NEW_CMD(VideoThreadCommand, ClearVRAM, sizeof(VideoThreadCommand)); }
VideoThreadCommand* GPUBackend::NewClearDisplayCommand() { return NEW_CMD(VideoThreadCommand, ClearDisplay, sizeof(VideoThreadCommand)); }
GPUBackendUpdateDisplayCommand* GPUBackend::NewUpdateDisplayCommand()
{
  return NEW_CMD(GPUBackendUpdateDisplayCommand, UpdateDisplay, sizeof(GPUBackendUpdateDisplayCommand));
}
GPUBackendSubmitFrameCommand* GPUBackend::NewSubmitFrameCommand()
{
  return NEW_CMD(GPUBackendSubmitFrameCommand, SubmitFrame, sizeof(GPUBackendUpdateDisplayCommand));
}
VideoThreadCommand* GPUBackend::NewClearCacheCommand() { return NEW_CMD(VideoThreadCommand, ClearCache, sizeof(VideoThreadCommand)); }
VideoThreadCommand* GPUBackend::NewBufferSwappedCommand() { return NEW_CMD(VideoThreadCommand, BufferSwapped, sizeof(VideoThreadCommand)); }
GPUBackendReadVRAMCommand* GPUBackend::NewReadVRAMCommand()
{
  return NEW_CMD(GPUBackendReadVRAMCommand, ReadVRAM, sizeof(GPUBackendReadVRAMCommand));
}
GPUBackendFillVRAMCommand* GPUBackend::NewFillVRAMCommand()
{
  return NEW_CMD(GPUBackendFillVRAMCommand, FillVRAM, sizeof(GPUBackendFillVRAMCommand));
}
GPUBackendUpdateVRAMCommand* GPUBackend::NewUpdateVRAMCommand(u32 num_words)
{
  return NEW_CMD(GPUBackendUpdateVRAMCommand, UpdateVRAM, sizeof(GPUBackendUpdateVRAMCommand) + num_words * static_cast<u32>(sizeof(u16)));
}
GPUBackendCopyVRAMCommand* GPUBackend::NewCopyVRAMCommand()
{
  return NEW_CMD(GPUBackendCopyVRAMCommand, CopyVRAM, sizeof(GPUBackendCopyVRAMCommand));
}
GPUBackendSetDrawingAreaCommand* GPUBackend::NewSetDrawingAreaCommand()
{
  return NEW_CMD(GPUBackendSetDrawingAreaCommand, SetDrawingArea, sizeof(GPUBackendSetDrawingAreaCommand));
}
GPUBackendUpdateCLUTCommand* GPUBackend::NewUpdateCLUTCommand()
{
  return NEW_CMD(GPUBackendUpdateCLUTCommand, UpdateCLUT, sizeof(GPUBackendUpdateCLUTCommand));
}
GPUBackendDrawPolygonCommand* GPUBackend::NewDrawPolygonCommand(u32 num_vertices)
{
  auto* cmd = NEW_CMD(GPUBackendDrawPolygonCommand, DrawPolygon,
                      static_cast<u32>(sizeof(GPUBackendDrawPolygonCommand) + num_vertices * sizeof(GPUBackendDrawPolygonCommand::Vertex)));
  cmd->num_vertices = static_cast<u16>(num_vertices);
  return cmd;
}
GPUBackendDrawPrecisePolygonCommand* GPUBackend::NewDrawPrecisePolygonCommand(u32 num_vertices)
{
  auto* cmd = NEW_CMD(GPUBackendDrawPrecisePolygonCommand, DrawPrecisePolygon,
                      static_cast<u32>(sizeof(GPUBackendDrawPrecisePolygonCommand) +
                                       num_vertices * sizeof(GPUBackendDrawPrecisePolygonCommand::Vertex)));
  cmd->num_vertices = static_cast<u16>(num_vertices);
  return cmd;
}
GPUBackendDrawRectangleCommand* GPUBackend::NewDrawRectangleCommand()
{
  return NEW_CMD(GPUBackendDrawRectangleCommand, DrawRectangle, sizeof(GPUBackendDrawRectangleCommand));
}
GPUBackendDrawLineCommand* GPUBackend::NewDrawLineCommand(u32 num_vertices)
{
  auto* cmd = NEW_CMD(GPUBackendDrawLineCommand, DrawLine,
                      static_cast<u32>(sizeof(GPUBackendDrawLineCommand) + num_vertices * sizeof(GPUBackendDrawLineCommand::Vertex)));
  cmd->num_vertices = static_cast<u16>(num_vertices);
  return cmd;
}
GPUBackendDrawPreciseLineCommand* GPUBackend::NewDrawPreciseLineCommand(u32 num_vertices)
{
  auto* cmd = NEW_CMD(GPUBackendDrawPreciseLineCommand, DrawPreciseLine,
                      static_cast<u32>(sizeof(GPUBackendDrawPreciseLineCommand) +
                                       num_vertices * sizeof(GPUBackendDrawPreciseLineCommand::Vertex)));
  cmd->num_vertices = static_cast<u16>(num_vertices);
  return cmd;
}
#undef NEW_CMD
void GPUBackend::PushCommand(VideoThreadCommand* cmd) { VideoThread::PushCommand(cmd); }
void GPUBackend::PushCommandAndSync(VideoThreadCommand* cmd, bool) { VideoThread::PushCommand(cmd); }
bool GPUBackend::BeginQueueFrame() { return false; }
void GPUBackend::WaitForOneQueuedFrame() {}
bool GPUBackend::IsUsingHardwareBackend() { return true; } // a device backend: VRAM->CPU reads emit ReadVRAM records (gpu.cpp:2413-2430)

// ---------------------------------------------------------------------------------------------- subsystems the decoder pokes
TimingEvent::TimingEvent(const char* name, TickCount period, TickCount interval, TimingEventCallback callback, void* callback_param)
  : m_callback(callback), m_callback_param(callback_param), m_period(period), m_interval(interval), m_name(name)
{
}
TimingEvent::~TimingEvent() = default;
void TimingEvent::Activate() { m_active = true; }
void TimingEvent::Deactivate() { m_active = false; }
void TimingEvent::Schedule(TickCount) { m_active = true; }
void TimingEvent::SetIntervalAndSchedule(TickCount ticks) { m_interval = ticks; m_active = true; }
void TimingEvent::InvokeEarly(bool) {}
TickCount TimingEvent::GetTicksSinceLastExecution() const { return 0; }
GlobalTicks TimingEvents::GetGlobalTickCounter() { return s_global_ticks; }
void TimingEvents::SetGlobalTickCounter(GlobalTicks ticks) { s_global_ticks = ticks; }
bool TimingEvents::IsRunningEvents() { return false; }

void DMA::SetRequest(Channel, bool) {}
void InterruptController::SetLineState(IRQ, bool) {}
void Timers::AddTicks(u32, TickCount) {}
TickCount Timers::GetTicksUntilIRQ(u32) { return std::numeric_limits<TickCount>::max(); }
bool Timers::IsExternalIRQEnabled(u32) { return false; }
bool Timers::IsSyncEnabled(u32) { return false; }
bool Timers::IsUsingExternalClock(u32) { return false; }
void Timers::SetGate(u32, bool) {}

bool System::IsReplayingGPUDump() { return s_replaying; }
bool System::IsPALRegion() { return s_pal; }
bool System::IsRunaheadActive() { return false; }
void System::FrameDone() {}
u32 System::GetFrameNumber() { return s_frame_number; }
void System::IncrementFrameNumber() { s_frame_number++; }
void System::IncrementInternalFrameNumber() {}
GlobalTicks System::GetGlobalTickCounter() { return s_global_ticks; }
TickCount System::GetTicksPerSecond() { return 33868800; }
void System::SetVideoFrameRate(float) {}
void System::UpdateAutomaticResolutionScale() {}
void System::UpdateGTEAspectRatio() {}
const std::string& System::GetGameSerial()
{
  static const std::string s;
  return s;
}
bool System::GetFramePresentationParameters(GPUBackendFramePresentationParameters* frame)
{
  std::memset(static_cast<void*>(frame), 0, sizeof(*frame));
  return true;
}
float PerformanceCounters::GetFPS() { return 0.0f; }
float PerformanceCounters::GetVPS() { return 0.0f; }
void Host::QueueAsyncTask(std::function<void()>) {}
bool CPU::PGXP::GetPreciseVertex(u32, u32, int, int, int, int, float*, float*, float*) { return false; }
// The texture cache's tail of the GPU section for a cache that tracks nothing (gpu_hw_texture_cache.cpp:661-760): marker,
// zero VRAM-write records, and from version 86 the index of the last write (none).
bool GPUTextureCache::DoState(StateWrapper& sw, bool)
{
  if (sw.GetVersion() < 73)
    return true;
  if (!sw.DoMarker("GPUTextureCache"))
    return false;
  u32 num_vram_writes = 0, last_vram_write_index = 0xFFFFFFFFu;
  sw.Do(&num_vram_writes);
  sw.DoEx(&last_vram_write_index, 86, 0xFFFFFFFFu);
  return num_vram_writes == 0 && !sw.HasError(); // (states made by this harness only)
}
bool GPUTextureCache::GetStateSize(StateWrapper& sw, u32* size)
{
  *size = 0;
  if (sw.GetVersion() < 73)
    return true;
  const size_t start = sw.GetPosition();
  if (!sw.DoMarker("GPUTextureCache"))
    return false;
  u32 num_vram_writes = 0, last_vram_write_index = 0;
  sw.Do(&num_vram_writes);
  sw.DoEx(&last_vram_write_index, 86, 0xFFFFFFFFu);
  // record sizes: gpu_hw_texture_cache.cpp:62-64 (two GSVector4i + 64-bit hash + u32; rect + 2 u32 + hash + 256 u16)
  constexpr u32 vram_write_size = 16 * 2 + 8 + 4, palette_record_size = 16 + 4 + 4 + 8 + 2 * 256;
  for (u32 i = 0; i < num_vram_writes && !sw.HasError(); i++)
  {
    sw.SkipBytes((sw.GetVersion() < 86) ? (vram_write_size - 4) : vram_write_size);
    u32 num_palette_records = 0;
    sw.Do(&num_palette_records);
    sw.SkipBytes(static_cast<size_t>(num_palette_records) * palette_record_size);
  }
  if (sw.HasError())
    return false;
  *size = static_cast<u32>(sw.GetPosition() - start);
  return true;
}
const char* Settings::GetDisplayCropModeName(DisplayCropMode) { return ""; }
std::string EmuFolders::DataRoot;

// dump recorder: never instantiated here
GPUDump::Recorder::~Recorder() = default;
bool GPUDump::Recorder::Close(Error*) { return true; }
bool GPUDump::Recorder::IsFinished() { return true; }
void GPUDump::Recorder::BeginGP0Packet(u32) {}
void GPUDump::Recorder::EndGP0Packet() {}
void GPUDump::Recorder::WriteGP0Packet(u32) {}
void GPUDump::Recorder::WriteGP1Packet(u32) {}
void GPUDump::Recorder::WriteGP1Command(GP1Command, u32) {}
void GPUDump::Recorder::WriteVSync(u64) {}
void GPUDump::Recorder::WriteWords(const u32*, size_t) {}
void GPUDump::Recorder::WriteDiscardVRAMRead(u32, u32) {}

// ImGui / images / errors / logging
bool ImGui::CollapsingHeader(const char*, ImGuiTreeNodeFlags) { return false; }
void ImGui::Text(const char*, ...) {}
Image::Image(u32, u32, ImageFormat) {}
bool Image::SaveToFile(const char*, u8, Error*) const { return false; }
Error::Error() = default;
Error::~Error() = default;
void Error::SetStringFmtArgs(fmt::string_view, fmt::format_args) {}
Log::Level Log::GetLogLevel() { return Log::Level::None; }
void Log::WriteFmtArgs(Log::MessageCategory, fmt::string_view, fmt::format_args) {}
void Log::WriteFuncNameFmtArgs(Log::MessageCategory, const char*, fmt::string_view, fmt::format_args) {}
int StringUtil::Strncasecmp(const char* a, const char* b, size_t n) { return strncasecmp(a, b, n); }
[[noreturn]] void Y_OnPanicReached(const char* msg, const char* func, const char* file, unsigned line)
{
  std::fprintf(stderr, "reference panic: %s (%s, %s:%u)\n", msg, func, file, line);
  std::abort();
}
void Y_OnAssertFailed(const char* msg, const char* func, const char* file, unsigned line)
{
  std::fprintf(stderr, "reference assertion: %s (%s, %s:%u)\n", msg, func, file, line);
  std::abort();
}

// paths only reached from the screenshot / dump-recording commands of the UI
TinyString Core::GetTinyStringSettingValue(const char*, const char*, std::string_view) { return TinyString(); }
FileSystem::AtomicRenamedFileDeleter::~AtomicRenamedFileDeleter() = default;
void FileSystem::AtomicRenamedFileDeleter::operator()(std::FILE*) {}
bool FileSystem::WriteAtomicRenamedFile(std::string, const void*, size_t, Error*) { return false; }
void GPUBackend::RenderScreenshotToFile(const std::string_view, DisplayScreenshotMode, u8, bool) {}
std::unique_ptr<GPUDump::Recorder> GPUDump::Recorder::Create(std::string, std::string_view, u32, Error*) { return {}; }
bool GPUDump::Recorder::Compress(const std::string&, GPUDumpCompressionMode, Error*) { return false; }
void Host::AddIconOSDMessage(OSDMessageType, std::string, const char*, std::string) {}
std::string_view Host::TranslateToStringView(std::string_view, std::string_view msg, std::string_view) { return msg; }
void Log::WriteFuncName(Log::MessageCategory, const char*, std::string_view) {}
std::string_view Path::GetExtension(std::string_view) { return {}; }
std::string_view Path::GetFileName(std::string_view path) { return path; }
std::string Path::ReplaceExtension(std::string_view path, std::string_view) { return std::string(path); }
std::optional<GPUDumpCompressionMode> Settings::ParseGPUDumpCompressionMode(std::string_view) { return std::nullopt; }
bool StringUtil::EqualNoCase(std::string_view a, std::string_view b)
{
  return a.size() == b.size() && strncasecmp(a.data(), b.data(), a.size()) == 0;
}
WindowInfo::WindowInfo() = default;

// g_settings / g_bus: zero-filled storage under the reference's names (the Settings constructor lives in settings.cpp,
// which needs the whole front end); the handful of fields the decoder reads are set by psxdec_init.
alignas(64) unsigned char psxdec_settings_storage[sizeof(Settings)] asm("g_settings");
alignas(64) unsigned char psxdec_bus_storage[sizeof(Bus::Globals)] asm("g_bus");

// ---------------------------------------------------------------------------------------------- C entry points
extern "C" {

// modulation / crop / timing settings the decoder reads.  crop_mode: DisplayCropMode value; pal: console region.
void psxdec_init(int pal, int crop_mode, int force_progressive)
{
  Settings& s = g_settings;
  s.gpu_fifo_size = 16;           // settings.h DEFAULT_GPU_FIFO_SIZE
  s.gpu_max_run_ahead = 128;      // settings.h DEFAULT_GPU_MAX_RUN_AHEAD
  s.display_crop_mode = static_cast<DisplayCropMode>(crop_mode);
  s.display_deinterlacing_mode = force_progressive ? DisplayDeinterlacingMode::Progressive : DisplayDeinterlacingMode::Disabled;
  s.gpu_resolution_scale = 1;
  s_pal = pal != 0;
  s_replaying = true;
  s_frame_number = 0;
  s_global_ticks = 0;
  s_arena.clear();
  GPU::Initialize();
  GPU::Reset(true);
  s_arena.clear();
}

// type: GPUDump::PacketType (0x00 GP0 data, 0x01 GP1 data, 0x02 VSync)
void psxdec_packet(unsigned type, const unsigned* words, size_t n)
{
  GPU::ProcessGPUDumpPacket(static_cast<GPUDump::PacketType>(type), std::span<const u32>(words, n));
}

size_t psxdec_records_size(void) { return s_arena.size(); }
void psxdec_take_records(void* dst)
{
  std::memcpy(dst, s_arena.data(), s_arena.size());
  s_arena.clear();
}

// GPU::DoState through the real StateWrapper.  save: returns the number of bytes written (or 0); load: 1 on success.
size_t psxdec_save_state(void* dst, size_t cap, unsigned version)
{
  StateWrapper sw(std::span<u8>(static_cast<u8*>(dst), cap), StateWrapper::Mode::Write, version);
  if (!GPU::DoState(sw) || sw.HasError())
    return 0;
  return sw.GetPosition();
}
int psxdec_load_state(const void* src, size_t bytes, unsigned version)
{
  StateWrapper sw(std::span<const u8>(static_cast<const u8*>(src), bytes), StateWrapper::Mode::Read, version);
  s_arena.clear();
  const bool ok = GPU::DoState(sw) && !sw.HasError();
  return ok ? 1 : 0;
}
unsigned psxdec_read_gpustat(void) { return GPU::ReadRegister(4); }
unsigned psxdec_state_version(void);

} // extern "C"
