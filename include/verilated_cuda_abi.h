/* -*- mode: C -*- */
/*************************************************************************
 *
 * verilated_cuda_abi.h -- the C ABI of the B200 batched evaluator.
 *
 * This is the drop-in boundary: a thin `extern "C"` surface
 * with plain pointers and sizes that replaces, for N instances at once, what a
 * testbench does with the generated C++ model class:
 *
 *   reference (file:line under /root/reference)              here
 *   ------------------------------------------------------  ---------------------------
 *   new VerilatedContext; randReset(); randSeed()            vlcu_create(..., seed, rand_reset)
 *     include/verilated.h:371,432-433; verilated.cpp:3362,3691
 *   new Vtop{contextp, "TOP"}                                vlcu_create(design, n, ...)
 *     src/V3EmitCModel.cpp:159-161,275-359
 *   top->port = value   (public port references)             vlcu_set_input / vlcu_port_info
 *     src/V3EmitCModel.cpp:108-117,289-296
 *   value = top->port                                         vlcu_get_output
 *   contextp->timeInc(n)      include/verilated.h:568         vlcu_time_inc
 *   top->eval()               src/V3EmitCModel.cpp:171-183,403-437   vlcu_eval
 *   while (!gotFinish) { timeInc; poke; eval; }               vlcu_run (K cycles fused on device)
 *     examples/make_tracing_c/sim_main.cpp:71-106
 *   contextp->gotFinish()/gotError()  verilated.h:560-566     vlcu_status
 *   $display/$write output    include/verilated.cpp (VL_WRITEF_NX)   vlcu_messages
 *   top->final()              src/V3EmitCModel.cpp:185-203    vlcu_final
 *   delete top                                                vlcu_destroy
 *   contextp->coveragep()->write(file)                        vlcu_coverage_reduce / _write
 *     include/verilated_cov.cpp:368-438
 *   tfp->dump(time)           include/verilated_trace.h:323   vlcu_run trace samples + vlcu_snapshot
 *   --lib-create C wrapper precedent: src/V3ProtectLib.cpp:344-415
 *
 * Error model: every call returns 0 on success or a negative VLCU_E_* code;
 * vlcu_last_error() gives text.  $finish/$stop/fatal inside an instance never
 * terminate the process, they are recorded in that instance's status.
 * There is NO CPU fallback: without a CUDA device vlcu_create() fails.
 *
 *************************************************************************/
#ifndef VERILATOR_VERILATED_CUDA_ABI_H_
#define VERILATOR_VERILATED_CUDA_ABI_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VLCU_ABI_VERSION 3

enum {
    VLCU_OK = 0,
    VLCU_E_ARG = -1, /* bad argument (unknown design/port/testbench, null pointer, range) */
    VLCU_E_CUDA = -2, /* CUDA runtime error (text in vlcu_last_error) */
    VLCU_E_NODEVICE = -3, /* no usable CUDA device: this library has no CPU path */
    VLCU_E_NCCL = -4, /* NCCL missing or failed */
    VLCU_E_STATE = -5, /* call not valid in the model's current state */
    VLCU_E_IO = -6 /* file could not be written */
};

typedef struct vlcu_model vlcu_model_t; /* opaque: owns all device memory of one shard */

enum { VLCU_DIR_IN = 0, VLCU_DIR_OUT = 1, VLCU_DIR_INOUT = 2 };

typedef struct {
    const char* name; /* port name, as the public member of Vtop */
    int dir; /* VLCU_DIR_* */
    int width; /* bits */
    int elem_bytes; /* C storage of one element: 1 CData, 2 SData, 4 IData/EData, 8 QData */
    int elems; /* number of elements (VlWide words), 1 for <= 64 bit ports */
} vlcu_port_info_t;

/* Per-instance status, the batched VerilatedContext flags */
#define VLCU_STATUS_DIDINIT 0x01u
#define VLCU_STATUS_FINISH 0x02u /* gotFinish() */
#define VLCU_STATUS_ERROR 0x04u /* gotError() */
#define VLCU_STATUS_STOPPED 0x08u /* $stop / assertion failure reached vl_stop */
#define VLCU_STATUS_FATAL 0x10u /* fatal, see fatal_code */
#define VLCU_STATUS_MSGDROP 0x20u /* message log overflowed */
typedef struct {
    uint64_t time; /* VerilatedContext::time() */
    uint64_t first_finish_time, first_stop_time, first_fatal_time;
    uint32_t status; /* VLCU_STATUS_* */
    uint32_t error_count;
    uint32_t finish_count, stop_count, fatal_count;
    uint32_t first_stop_line, first_fatal_line, fatal_code;
    uint32_t msg_count;
    uint32_t pad_;
} vlcu_status_t;

typedef struct {
    uint64_t time;
    uint32_t id; /* index for vlcu_message_format() */
    uint32_t nargs;
    uint64_t args[8];
} vlcu_msg_t;

/* vlcu_run flags */
#define VLCU_RUNF_TB_BEGIN 0x1u /* apply the testbench's initial port values first */
#define VLCU_RUNF_STOP_ON_FINISH 0x2u /* per instance: stop evaluating after gotFinish() */
#define VLCU_RUNF_STIM_DEVICE 0x4u /* stim points to device memory */
#define VLCU_RUNF_OBS_DEVICE 0x8u /* obs points to device memory */
#define VLCU_RUNF_STIM_BROADCAST 0x10u /* one stimulus row serves all instances */
#define VLCU_RUNF_TRACE_CHANGES 0x80u /* sampled instances record change streams (below) instead of full records */
#define VLCU_RUNF_BULK_1D 0x40u /* "trace": stage rows with per-word 1-D bulk copies instead of 2-D tensor tiles */
#define VLCU_RUNF_NO_STREAM 0x20u /* "trace": read/write rows with plain global accesses instead of TMA-staged stages */

typedef struct {
    const char* tb; /* testbench: "none", "trace", or a design-specific one */
    uint64_t n_evals; /* eval() calls; one clock period is two */
    uint32_t flags; /* VLCU_RUNF_* */
    uint32_t pad_;
    /* "trace" testbench: uint32 [eval][in_word][instance (padded stride, vlcu_padded_instances)] */
    const uint32_t* stim;
    /* optional capture of every output port after each of the first obs_evals evals:
       uint32 [eval][out_word][instance (padded stride)] */
    uint32_t* obs;
    uint64_t obs_evals;
    /* optional waveform sampling: full state records of a few instances after each eval,
       host buffer of trace_evals * n_samples * vlcu_trace_record_bytes() bytes */
    const uint64_t* sample_instances;
    uint32_t n_samples;
    /* VLCU_RUNF_TRACE_CHANGES: trace_out holds n_samples streams of trace_stream_words uint32
       each (vlcu_vcd_write_changes documents the format); 0 otherwise */
    uint32_t trace_stream_words;
    void* trace_out;
    uint64_t trace_evals;
    /* parameters interpreted by the design-specific testbench; the built-in "fuzz" testbench
       (random stimulus made in the kernel, include/verilated_cuda_launch.h VlcuTbFuzz) takes
       tb_seed, tb_flags = VLCU_TBF_FUZZ_*, tb_arg[0] = clock periods of reset, tb_arg[1] = bit
       mask of the input row words that are sparse (non-zero in 1 period of 32) */
    uint64_t tb_seed;
    uint32_t tb_flags;
    uint32_t tb_arg[4];
    uint32_t pad3_;
    /* optional per-run data of a design-specific testbench (ABI 3): tb_data_bytes bytes copied
       host -> device before the launch (asynchronously when page-locked) and handed to the
       testbench as its first pointer argument -- e.g. one 64-bit program seed per instance for
       the RV32I fuzz campaign ("rv32_seeds"), 8 bytes per instance instead of a 4 KB image.
       With VLCU_RUNF_TBDATA_DEVICE tb_data already is device memory. */
    const void* tb_data;
    uint64_t tb_data_bytes;
} vlcu_run_desc_t;
#define VLCU_RUNF_TBDATA_DEVICE 0x100u
#define VLCU_TBF_FUZZ_RESET_LOW 0x1u  /* input row word 1 is an active-low reset */
#define VLCU_TBF_FUZZ_RESET_HIGH 0x2u /* ... an active-high reset */
#define VLCU_TBF_FUZZ_UNIFORM 0x4u    /* all instances see the same stimulus */

/* ---- library ---------------------------------------------------------- */
int vlcu_abi_version(void);
const char* vlcu_last_error(void);
int vlcu_device_count(void);
int vlcu_design_count(void);
const char* vlcu_design_name(int index);

/* Page-locked host memory for the buffers handed to vlcu_set_input / vlcu_get_output / vlcu_run
   (plain malloc'ed memory works too, pinned memory makes the copies asynchronous DMA) */
void* vlcu_host_alloc(size_t bytes);
void vlcu_host_free(void* ptr);

/* Device buffers for stimulus / observation rows that stay resident in HBM between runs
   (VLCU_RUNF_STIM_DEVICE / VLCU_RUNF_OBS_DEVICE) */
void* vlcu_device_alloc(vlcu_model_t* h, size_t bytes);
void vlcu_device_free(vlcu_model_t* h, void* ptr);
int vlcu_device_copy(vlcu_model_t* h, void* dst, const void* src, size_t bytes, int to_device);
/* vlcu_run with host-resident "trace" stimulus larger than two chunks streams it through the
   device in eval-chunks of about `bytes` (default 64 MB; 0 = always stage the whole run).  The
   copy-in of chunk c+1, the kernel of chunk c and the copy-out of chunk c-1 overlap ONLY when
   desc->stim and desc->obs are page-locked (vlcu_host_alloc): with pageable memory every
   cudaMemcpyAsync is synchronous and the pipeline degrades to copy, run, copy. */
int vlcu_set_pipeline_chunk(vlcu_model_t* h, size_t bytes);

/* ---- model life cycle -------------------------------------------------- */
/* n_instances models of `design` on CUDA device `device`; instance i of this shard has global
   index inst_base+i and PRNG seed (seed_base ^ global index); rand_reset as randReset();
   msg_cap = $display records kept per instance; plusargs = space separated "+arg" list. */
vlcu_model_t* vlcu_create(const char* design, uint64_t n_instances, int device, uint64_t inst_base,
                          uint64_t seed_base, int rand_reset, uint32_t msg_cap,
                          const char* plusargs);
void vlcu_destroy(vlcu_model_t* h);
int vlcu_reset(vlcu_model_t* h); /* re-run construction (ctor_var_reset), time = 0 */

uint64_t vlcu_instances(const vlcu_model_t* h);
uint64_t vlcu_padded_instances(const vlcu_model_t* h); /* row stride of stim/obs buffers */
uint32_t vlcu_state_bytes(const vlcu_model_t* h); /* S_persist, bytes per instance */
uint32_t vlcu_in_words(const vlcu_model_t* h);
uint32_t vlcu_out_words(const vlcu_model_t* h);
uint32_t vlcu_in_bytes(const vlcu_model_t* h); /* S_in  (C storage bytes of input ports) */
uint32_t vlcu_out_bytes(const vlcu_model_t* h); /* S_out */
uint32_t vlcu_trace_record_bytes(const vlcu_model_t* h);
/* Persistent bytes one triggered eval reads / writes (S_state_rd, S_state_wr of the roofline
   model in DESIGN.md); equal to vlcu_state_bytes() when the design does not say */
uint32_t vlcu_touched_read_bytes(const vlcu_model_t* h);
uint32_t vlcu_touched_write_bytes(const vlcu_model_t* h);

/* ---- ports ---------------------------------------------------------------- */
int vlcu_port_count(const vlcu_model_t* h);
int vlcu_port_info(const vlcu_model_t* h, int port, vlcu_port_info_t* out);
int vlcu_port_index(const vlcu_model_t* h, const char* name);
/* Host array of per-instance values, element i at src + i*stride_bytes, each
   elems*elem_bytes long (the C type the reference model exposes); stride 0 = broadcast.
   Values are masked to the port width on ingest. */
int vlcu_set_input(vlcu_model_t* h, int port, const void* src, size_t stride_bytes);
int vlcu_get_output(vlcu_model_t* h, int port, void* dst, size_t stride_bytes);
/* Any member by name (ports, internal signals, memories): same host layout */
int vlcu_member_count(const vlcu_model_t* h);
const char* vlcu_member_name(const vlcu_model_t* h, int member);
int vlcu_member_index(const vlcu_model_t* h, const char* name);
int vlcu_member_info(const vlcu_model_t* h, int member, vlcu_port_info_t* out);
int vlcu_set_member(vlcu_model_t* h, int member, const void* src, size_t stride_bytes);
int vlcu_get_member(vlcu_model_t* h, int member, void* dst, size_t stride_bytes);
/* Asynchronous ingest for pipelined campaigns: vlcu_stage_member starts the host->device copy of
 * a member's NEXT value (src in page-locked memory, see vlcu_host_alloc) on a copy stream into a
 * staging buffer while a vlcu_run is still computing; vlcu_commit_staged enqueues, on the
 * model's stream, the move of everything staged into the state.  Neither call blocks the host;
 * src must stay unchanged until the first synchronising call after the commit.  mask_width != 0
 * masks to the declared width as vlcu_set_input does.  No reference counterpart: it is
 * "top->port = value" for the next batch, issued early. */
int vlcu_stage_member(vlcu_model_t* h, int member, const void* src, size_t stride_bytes,
                      int mask_width);
int vlcu_commit_staged(vlcu_model_t* h);
/* Raw device pointer + padded instance stride of a member's SoA array (a tiled design, see
   vlcu_tile(): lane 0 of tile 0, *npad = the tile size) */
int vlcu_member_device_ptr(vlcu_model_t* h, int member, void** ptr, uint64_t* npad);

/* ---- evaluation ------------------------------------------------------------- */
int vlcu_time_inc(vlcu_model_t* h, uint64_t add); /* all instances */
int vlcu_eval(vlcu_model_t* h); /* one eval() per instance */
int vlcu_run(vlcu_model_t* h, const vlcu_run_desc_t* desc); /* n_evals fused in one kernel */
int vlcu_final(vlcu_model_t* h);
int vlcu_sync(vlcu_model_t* h);
/* Benchmark hygiene: overwrite a scratch buffer larger than the 126 MB L2 on the model's stream */
int vlcu_flush_l2(vlcu_model_t* h);
/* Device time of the last vlcu_eval/vlcu_run kernel (CUDA events on the model's stream) */
int vlcu_last_kernel_ms(vlcu_model_t* h, float* ms);
uint64_t vlcu_launch_count(const vlcu_model_t* h); /* kernels launched so far by this model */

/* ---- observation ------------------------------------------------------------- */
int vlcu_status(vlcu_model_t* h, vlcu_status_t* out /*[n_instances]*/);
int vlcu_signatures(vlcu_model_t* h, uint64_t* out /*[n_instances]*/);
int vlcu_reset_signatures(vlcu_model_t* h);
/* The same read-backs without stalling the host: signatures, status records and the listed
   members (host layout of vlcu_get_member, tightly packed) as they are once the work issued so far
   has run, copied aside on the device and sent to the host on a separate stream.  Returns at once;
   launch the next vlcu_run, then vlcu_results_wait() before reading the destinations (page-locked
   memory, see vlcu_host_alloc).  sigs_out / status_out may be NULL.  One read-back at a time. */
int vlcu_results_begin(vlcu_model_t* h, uint64_t* sigs_out, vlcu_status_t* status_out, int n_members,
                       const int* members, void* const* member_dsts);
int vlcu_results_wait(vlcu_model_t* h);
int vlcu_messages(vlcu_model_t* h, uint64_t instance, vlcu_msg_t* out, uint32_t cap,
                  uint32_t* count);
const char* vlcu_message_format(const vlcu_model_t* h, uint32_t id);
/* All persistent members of one instance, packed at their vlcu_member_offset() */
int vlcu_snapshot(vlcu_model_t* h, uint64_t instance, void* out, size_t out_bytes);
uint32_t vlcu_member_offset(const vlcu_model_t* h, int member);
/* Instances per tile of the model image: 0 for a structure-of-arrays design (every member is one
   [element][instance] array over the padded shard), else the design was printed in the tiled
   form (include/verilated_cuda.h "Tiled model image") and its state block is an array of tiles of
   vlcu_state_bytes() * vlcu_tile() bytes.  Only vlcu_member_device_ptr() users need to know. */
uint32_t vlcu_tile(const vlcu_model_t* h);

/* One instance as a verilated_save image (include/verilated_save.h:193-223): the file
   `VerilatedSave os; os.open(f); os << *top;` writes for a --savable build of the same design,
   so a failing instance of a campaign can be loaded into the ordinary C++ model with
   `VerilatedRestore os; os.open(f); os >> *top;` and replayed under a debugger; and the reverse,
   a state saved by the C++ model loaded into instance `instance`.  The image's randSeed /
   randReset must be the instance's own.  Not available for --coverage builds (the reference
   refuses --savable with --coverage). */
int vlcu_save_instance(vlcu_model_t* h, uint64_t instance, const char* filename);
int vlcu_restore_instance(vlcu_model_t* h, uint64_t instance, const char* filename);
/* The bytes of a freshly constructed VerilatedContext's serialized part as this library lays
   them out (returns their count); for checking against the runtime the images are loaded by */
int vlcu_save_context_image(void* out, size_t out_bytes);

/* Waveform of ONE sampled instance as VCD text, written exactly as VerilatedVcdC would write it
   for that model instance with a dump() after every eval (include/verilated_vcd_c.cpp:117-129,
   191-203,317-470,620-690; include/verilated_trace_imp.h): `trace` is the buffer vlcu_run filled
   (trace_evals x n_samples records), `sample` selects the instance's slot in it. */
int vlcu_vcd_write(const vlcu_model_t* h, const void* trace, uint64_t trace_evals,
                   uint32_t n_samples, uint32_t sample, const char* filename);

/* The same from a change stream (VLCU_RUNF_TRACE_CHANGES): the device compares every trace word
   with its value at the previous dump and records only the differences, so the device->host copy
   and the host-side work scale with signal activity, not with design size x evals (the role of
   the chg* functions of include/verilated_trace.h:384-432).  Stream of one sample, uint32 words:
   [0] words used  [1] dumps dropped for lack of room  [2] dumps recorded  [3] reserved, then per
   dump: [0x80000000 | n] [time lo] [time hi] and n x { [trace word index] [value] }; the first
   dump lists every word.  `streams` is the buffer vlcu_run filled. */
int vlcu_vcd_write_changes(const vlcu_model_t* h, const uint32_t* streams, uint32_t stream_words,
                           uint32_t n_samples, uint32_t sample, const char* filename);

/* ---- coverage ------------------------------------------------------------------ */
uint32_t vlcu_coverage_bins(const vlcu_model_t* h); /* counters per instance */
uint32_t vlcu_coverage_points(const vlcu_model_t* h); /* registered points (>= bins: points can share a counter) */
/* key fields of registration `point`, as passed to VL_COVER_INSERT (include/verilated_cov.h:70),
   and the counter (bin) it reads */
int vlcu_coverage_point(const vlcu_model_t* h, uint32_t point, uint32_t* bin, const char** filename,
                        int* lineno, int* column, const char** hier, const char** page,
                        const char** comment, const char** linescov);
/* Sum the per-instance uint32 counters of this shard on the device into uint64 per bin; when a
   communicator is attached (vlcu_comm_init) the per-shard sums are then all-reduced with NCCL. */
int vlcu_coverage_reduce(vlcu_model_t* h, uint64_t* out /*[bins]*/);
/* Write `counts` in the reference's coverage.dat format (include/verilated_cov.cpp:368-438) */
int vlcu_coverage_write(const vlcu_model_t* h, const uint64_t* counts, const char* filename);

/* ---- multi-GPU: one model shard per process/GPU, NCCL only for the coverage sum ---- */
#define VLCU_NCCL_ID_BYTES 128
int vlcu_comm_unique_id(void* id_out /*[VLCU_NCCL_ID_BYTES]*/);
int vlcu_comm_init(vlcu_model_t* h, const void* id, int rank, int world);
int vlcu_comm_destroy(vlcu_model_t* h);

#ifdef __cplusplus
}
#endif
#endif /* guard */
