// -*- mode: C++; c-file-style: "cc-mode" -*-
//*************************************************************************
//
// verilated_cuda_host.h -- host-side description of an emitted design and
// the interface between emitted design code and the batch runtime
// (verilated_cuda.cu).  This is the batched counterpart of what the
// generated Vtop__Syms / Vtop.h give the host runtime: the list of ports
// (V3EmitCModel.cpp:108-117), the member layout (V3EmitCHeaders.cpp:73-143),
// the coverage-point registrations (V3EmitCImp.cpp:179-262) and the
// $display format strings.
//
//*************************************************************************
#ifndef VERILATOR_VERILATED_CUDA_HOST_H_
#define VERILATOR_VERILATED_CUDA_HOST_H_

#include "verilated_cuda.h"

// Run flags (also used from the C ABI, include/verilated_cuda_abi.h)
enum : uint32_t {
    VLCU_RUN_TB_BEGIN = 1u << 0,  // call the testbench's begin() first (initial port values)
    VLCU_RUN_STOP_ON_FINISH = 1u << 1,  // an instance stops evaluating once gotFinish() is set
    VLCU_RUN_TRACE_CHANGES = 1u << 7,  // sampled instances record change streams, not full records
    VLCU_RUN_BULK_1D = 1u << 6,  // "trace" runs: per-word 1-D bulk copies instead of 2-D tensor tiles
    VLCU_RUN_NO_STREAM = 1u << 5,  // "trace" runs: direct global accesses instead of TMA-staged rows
};
#define VLCU_TRACE_HDR 16  // bytes before the state record in a trace sample: u64 time, u32 status, u32 valid
// A trace sample = [header][packed state, kStateBytes][trace words, kTraceWords x u32, padded to
// an even count so that every sample starts 8-byte aligned]
#define VLCU_TRACE_WORDS_PADDED(words) (((words) + 1U) & ~1U)

// TMA tensor maps (CUtensorMap, 128 bytes each) of a "trace" run's stimulus and observation
// rows seen as 2-D uint32 tensors [eval * words + word][instance]; box = {VLCU_BLOCK instances,
// all words of one eval}.  Encoded by the runtime per launch, passed to the kernel by value.
struct alignas(64) VlcuRowMaps final {
    uint64_t in[16];
    uint64_t out[16];
};

// Per-launch parameters of the cycle-loop kernel
struct VlcuRunDev final {
    uint64_t nEvals;  // eval() calls to make (two per clock period)
    uint64_t evalBase;  // index of the first eval (for resumed runs)
    uint64_t npad;  // padded instance count (row stride of stim/obs)
    uint32_t flags;
    uint32_t stimBroadcast;  // stimulus rows hold one value for all instances
    const uint32_t* stim;  // [eval][input word][instance] (VlcuTbTrace)
    uint32_t* obs;  // [eval][output word][instance] or null
    uint64_t obsEvals;
    uint64_t* sigs;  // [instance] running output signature
    uint64_t* evalsDone;  // [instance] evals executed (stop-on-finish accounting) or null
    // Waveform sampling of a few instances
    const uint64_t* sampleIds;  // local instance numbers
    uint32_t nSamples;
    uint8_t* trace;  // [eval][sample][VLCU_TRACE_HDR + stateBytes]
    uint64_t traceEvals;
    // Testbench parameters
    uint64_t tbSeed;
    uint32_t tbFlags;
    uint32_t tbArg[4];
    void* tbPtr[2];
    const VlcuRowMaps* rowMaps;  // HOST pointer (launch-time only) or null: no tensor maps
    // Waveform sampling as change streams (VLCU_RUN_TRACE_CHANGES): `trace` holds one stream of
    // traceCapWords uint32 per sample, traceScratch 2 x padded trace words per sample (the
    // current and the previously recorded value of every trace code)
    uint32_t traceCapWords;
    uint32_t* traceScratch;
};

// Change stream of one sampled instance (uint32 words):
//   [0] words used so far (header included)  [1] overflow: dumps dropped for lack of room
//   [2] dumps recorded                        [3] reserved
//   then per dump: [VLCU_TRACE_DUMP_MARK | n] [time lo] [time hi]  n x { [trace word index] [value] }
// The first dump of a run lists every trace word; later ones only the words whose value differs
// from the previous dump -- the device-side form of VerilatedTrace's chg* change detection
// (verilated_trace.h:384-432).
#define VLCU_TRACE_STREAM_HDR 4U
#define VLCU_TRACE_DUMP_MARK 0x80000000U

enum VlcuMemberKind : uint8_t {
    VLCU_MEMBER_IN = 0,  // primary input port
    VLCU_MEMBER_OUT = 1,  // primary output port
    VLCU_MEMBER_INOUT = 2,
    VLCU_MEMBER_STATE = 3,  // internal persistent signal / trigger prev / NBA shadow
    VLCU_MEMBER_COVERAGE = 4,  // uint32_t __Vcoverage[nBins]
    VLCU_MEMBER_MEMORY = 5,  // VlcuGlobalArray: instance-major, not interleaved
};

struct VlcuMemberDesc final {
    const char* name;  // C++ member name in Vtop___024root
    const char* hierName;  // Verilog hierarchical name (for tracing), or nullptr
    uint8_t kind;
    uint8_t elemBytes;  // 1 CData, 2 SData, 4 IData/EData, 8 QData
    uint16_t flags;  // VLCU_MEMBERF_*
    uint32_t elems;  // words of a VlWide, depth of an array, else 1
    uint32_t width;  // bits (of one array element)
    int32_t msb, lsb;  // declared range
    uint32_t soaOffset;  // SoA designs: byte offset per instance in the SoA block (the member starts at
                         // soaOffset * npad); tiled designs: byte offset of lane 0 inside a tile
};
enum : uint16_t {
    VLCU_MEMBERF_LANE_MAJOR = 1u << 0,  // the elements of one instance are contiguous ([lane][element])
};

struct VlcuCoverPoint final {  // arguments of __vlCoverInsert / __vlCoverToggleInsert
    uint32_t bin;  // index of the counter in __Vcoverage; several points may share one counter
                   // (V3CoverageJoin merges toggles of signals that are always equal)
    const char* filename;
    int lineno;
    int column;
    const char* hier;  // relative to the model's name, e.g. ".top.sub"
    const char* page;  // "v_line/sub", "v_toggle/top", "v_user/sub"
    const char* comment;  // "block", "if", "else", "count_c[3]:0->1" ...
    const char* linescov;
};

// One statement of the model's trace_init functions (what V3TraceDecl.cpp builds and
// V3EmitCImp.cpp:643-700,842-857 prints as VL_TRACE_DECL_* / VL_TRACE_PUSH_PREFIX /
// VL_TRACE_POP_PREFIX), flattened: dtype functions are expanded at their call sites.
enum VlcuTraceOp : uint8_t { VLCU_TRACE_SIG = 0, VLCU_TRACE_PUSH = 1, VLCU_TRACE_POP = 2 };
// VerilatedTracePrefixType (include/verilated_trace.h:50-60)
enum VlcuTracePrefix : int32_t {
    VLCU_PREFIX_ARRAY_PACKED = 0,
    VLCU_PREFIX_ARRAY_UNPACKED = 1,
    VLCU_PREFIX_ROOTIO_WRAPPER = 2,
    VLCU_PREFIX_SCOPE_MODULE = 3,
    VLCU_PREFIX_SCOPE_INTERFACE = 4,
    VLCU_PREFIX_STRUCT_PACKED = 5,
    VLCU_PREFIX_STRUCT_UNPACKED = 6,
    VLCU_PREFIX_UNION_PACKED = 7,
};
struct VlcuTraceDecl final {
    uint8_t op;  // VlcuTraceOp
    uint8_t bussed;  // SIG: declared with a [msb:lsb] range (DECL_BUS/QUAD/WIDE) vs DECL_BIT
    const char* name;  // signal or prefix name
    uint32_t code;  // SIG: trace code relative to the model's base = word index in a trace sample
    int32_t arg;  // SIG: width in bits;  PUSH: VlcuTracePrefix
    int32_t elem;  // SIG: index of an unpacked array element (declared as name[elem]), else -1
    int32_t msb, lsb;
};

#ifdef __cplusplus
#include <string>
#include <vector>
// Builder of the coverage-point table from the registrations V3EmitCuda prints, one call per
// __vlCoverInsert / __vlCoverToggleInsert of the C++ model (V3EmitCImp.cpp:179-262): a toggle
// registration expands to two bins per bit, "name[bit]:1->0" then ":0->1"
struct VlcuCoverTable final {
    std::vector<VlcuCoverPoint> points;
    std::vector<std::string> comments;  // owns the generated comment strings
    void other(int bin, const char* file, int line, int column, const char* hier, const char* page,
               const char* comment, const char* linescov) {
        points.push_back(VlcuCoverPoint{static_cast<uint32_t>(bin), file, line, column, hier, page,
                                        nullptr, linescov});
        comments.push_back(comment);
    }
    void toggle(int bin, int begin, int end, int ranged, const char* file, int line, int column,
                const char* hier, const char* page, const char* comment) {
        const int step = (end >= begin) ? 1 : -1;
        for (int i = begin; i != end + step; i += step) {
            for (int j = 0; j < 2; ++j) {
                std::string c = comment;
                if (ranged) c += '[' + std::to_string(i) + ']';
                c += j ? ":0->1" : ":1->0";
                points.push_back(VlcuCoverPoint{static_cast<uint32_t>(bin), file, line, column, hier,
                                                page, nullptr, ""});
                comments.push_back(c);
                ++bin;
            }
        }
    }
    void finish() {
        for (size_t i = 0; i < points.size(); ++i) points[i].comment = comments[i].c_str();
    }
};
#endif

struct VlcuCoverScope final {
    uint32_t tileOffset;  // of counter 0 inside a tile
    uint32_t bins;
};

struct VlcuDesignInfo final {
    const char* name;  // design name used with vlcu_create()
    const char* topName;  // Verilog top module
    const VlcuMemberDesc* members;
    uint32_t nMembers;
    uint32_t stateBytes;  // S_persist: bytes per instance in the SoA block
    uint32_t inWords, outWords;  // 32-bit words per stimulus / observation row
    uint32_t inBytes, outBytes;  // S_in / S_out in C storage bytes (SURVEY 8d B_alg)
    // Persistent bytes one triggered eval reads / writes (SURVEY 8d S_state_rd / S_state_wr,
    // "touched memory words only"); 0 = the whole state (stateBytes)
    uint32_t touchedReadBytes, touchedWriteBytes;
    const VlcuCoverPoint* coverPoints;  // registrations, in __Vconfigure order
    uint32_t nCoverPoints;
    uint32_t coverBins;  // counters per instance: uint32_t __Vcoverage[coverBins]
    uint32_t coverageSoaOffset;  // of __Vcoverage, if coverBins
    const char* const* messages;  // $display/$write format strings by id
    uint32_t nMessages;
    const char* const* plusargs;  // $test$plusargs strings by index
    uint32_t nPlusargs;
    const VlcuTraceDecl* traceDecls;  // trace_init, in declaration order (nullptr: not traced)
    uint32_t nTraceDecls;
    uint32_t traceWords;  // 32-bit words of one trace sample (the model's trace codes)
    const uint32_t* traceOrder;  // trace codes in the order the model's trace functions dump them
    uint32_t nTraceOrder;
    // verilated_save image of one instance (what a --savable build's __Vserialize writes,
    // V3EmitCImp.cpp:277-350): member indices in serialization order and the structure checksum
    const uint32_t* saveOrder;
    uint32_t nSaveOrder;
    uint32_t saveFlags;  // VLCU_SAVE_*
    uint64_t saveCheckval;  // __Vcheckval
    const char* saveUnsupported;  // first member this runtime cannot serialize, or nullptr
    // what the Syms constructor passes to VerilatedContext::timeunit()/timeprecision()
    // (V3EmitCSyms.cpp:1191-1199), as the context stores them: -power of ten; -1 = not set
    int saveTimeunit, saveTimeprecision;
    // 0: structure of arrays over the whole shard (member m, element e, instance i at
    //    soaOffset*npad + (e*npad + i)*elemBytes).  Else the model image is tiled: `tile` instances
    //    per tile of stateBytes*tile bytes (verilated_cuda.h "Tiled model image")
    uint32_t tile;
    // tiled designs built --coverage: every scope keeps its own counters in the tile, ONE uint64
    // per bin for the tile's instances together (VlcuTileCov); the shard's counter b is counter
    // (b - first) of the scope whose range holds b
    const struct VlcuCoverScope* coverScopes;
    uint32_t nCoverScopes;
};
enum : uint32_t {
    VLCU_SAVE_ACTIVITY = 1u << 0,  // Syms::__Vserialize writes __Vm_activity first (--trace builds)
};

#if defined(__CUDACC__) || defined(VLCU_HOST_RUNTIME)
#include <cuda_runtime.h>
class VlcuDesignOps {
public:
    virtual ~VlcuDesignOps() {}
    virtual const VlcuDesignInfo& info() const = 0;
    virtual cudaError_t launchReset(const VlcuModelDev& m, cudaStream_t stream) const = 0;
    virtual cudaError_t launchFinal(const VlcuModelDev& m, cudaStream_t stream) const = 0;
    virtual cudaError_t launchRun(const char* tb, const VlcuModelDev& m, const VlcuRunDev& r,
                                  cudaStream_t stream) const = 0;
    virtual int tbCount() const = 0;
    virtual const char* tbName(int i) const = 0;
};
// Implemented by the runtime; called from each design's static initialiser
int vlcu_register_design(VlcuDesignOps* opsp);
#endif

#endif  // guard
