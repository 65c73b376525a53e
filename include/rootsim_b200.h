/*
 * rootsim_b200.h -- C-ABI of the B200-native Time Warp engine (the drop-in boundary).
 *
 * One shared library is produced per model by `rootsim-cc` (lib<model>.so): the model's unmodified
 * C sources, compiled as __device__ callbacks, fused with the engine kernels (sm_100a).  The host
 * runtime (host/rs_main.c, plain C, ROOT-Sim compatible CLI) and any foreign-language binding talk
 * to it ONLY through the entry points below: plain pointers and sizes, no CUDA or torch types.
 * There is no CPU execution path behind this ABI: without a usable CUDA device rs_dev_init fails
 * with RS_ERR_NO_DEVICE.
 *
 * Each entry point names the reference interface it stands in for (paths relative to the
 * HPDCS/ROOT-Sim 2.0.0 tree).
 */
#ifndef ROOTSIM_B200_H
#define ROOTSIM_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RS_ABI_VERSION 3

/* error codes (reference: rootsim_error(fatal=true) -> exit(EXIT_FAILURE), src/core/core.c:220-250) */
enum {
	RS_OK = 0,
	RS_ERR_NO_DEVICE = 1,		/* no CUDA device / driver: the product has no CPU fallback        */
	RS_ERR_BAD_CONFIG = 2,
	RS_ERR_OOM = 3,
	RS_ERR_CUDA = 4,
	RS_ERR_MODEL = 5,		/* model raised a fatal condition, see rs_dev_error() / status bits  */
	RS_ERR_CAPACITY = 6		/* a device-side pool overflowed, see status bits                    */
};

/* device status bits (rs_stats_t.status); the first three are the reference's three send-time checks
 * (src/communication/communication.c:280-294) */
#define RS_ST_BAD_RECEIVER	0x0001u	/* receiver >= n_prc_tot: warned and dropped (not fatal)        */
#define RS_ST_PAST_EVENT	0x0002u	/* timestamp < now: fatal                                       */
#define RS_ST_RESERVED_TYPE	0x0004u	/* event type >= 65532: fatal                                   */
#define RS_ST_MODEL_EXIT	0x0008u	/* model called exit()/abort()                                  */
#define RS_ST_ARENA_FULL	0x0010u	/* per-LP state arena exhausted (malloc returned NULL)           */
#define RS_ST_POOL_FULL		0x0020u	/* pending-event pool exhausted                                 */
#define RS_ST_WINDOW_FULL	0x0040u	/* window/grouping buffer exhausted                             */
#define RS_ST_OUT_FULL		0x0080u	/* per-window output (sent-log) buffer exhausted                */
#define RS_ST_HORIZON		0x0100u	/* far list full: too many events beyond the calendar ring's horizon
					 * (n_buckets*bucket_width); such events normally wait on the far list */
#define RS_ST_LATE_FULL		0x0200u	/* too many in-window arrivals for one LP                       */
#define RS_ST_PAYLOAD		0x0400u	/* event payload larger than max_payload                        */
#define RS_ST_NO_EVENTS		0x0800u	/* pending set became empty before the horizon                  */
#define RS_ST_XCHG_FULL		0x1000u	/* inter-rank exchange slot exhausted (raise xchg_cap)          */
#define RS_ST_PEER_FATAL	0x2000u	/* another rank stopped with a fatal status (its own bits say why) */
#define RS_ST_FATAL_MASK	(RS_ST_PAST_EVENT | RS_ST_RESERVED_TYPE | RS_ST_MODEL_EXIT | RS_ST_ARENA_FULL | \
				 RS_ST_POOL_FULL | RS_ST_WINDOW_FULL | RS_ST_OUT_FULL | RS_ST_HORIZON | RS_ST_LATE_FULL | RS_ST_PAYLOAD | \
				 RS_ST_XCHG_FULL | RS_ST_PEER_FATAL)

/* engine flags */
#define RS_F_TRACE		0x1u	/* keep the per-LP committed-trace digest (count/hash/last ts)   */
#define RS_F_KTIME		0x2u	/* time every kernel launch with CUDA events (diagnostic runs)    */
#define RS_F_SERIAL_RUN		0x4u	/* execute a round's ordinary LPs after its chain LPs, on one stream (default: beside
					 * them on a second stream; RS_F_KTIME implies the serial order so that its per-kernel
					 * times add up to the step) */

/*
 * Run configuration.  Mirrors the fields of the reference's `rootsim_config` that matter on this
 * path (src/core/init.h:57-80, defaults src/core/init.c:315-340) plus the sizing of the device pools.
 * Zero in a sizing field selects the engine default.
 */
typedef struct rs_config {
	uint32_t abi_version;		/* RS_ABI_VERSION                                                  */
	uint32_t n_lps;			/* --lp: total LPs in the simulation (all ranks)                   */
	uint32_t lp_first;		/* first global LP id owned by this engine (block partition,       */
	uint32_t lp_count;		/*   src/core/core.c:275-300); lp_count==0 means all n_lps        */
	uint64_t master_seed;		/* ~/.rootsim/numerical.conf value (src/lib/numerical.c:326-391)   */
	double simulation_time;		/* --simulation-time; 0 = run until OnGVT says stop                */
	uint32_t ckpt_period;		/* --p: an LP takes a checkpoint at its first event of a window and then every
					 * ckpt_period events inside it (src/mm/state.c:55-130)            */
	uint32_t gvt_snapshot_cycles;	/* --gvt-snapshot-cycles: OnGVT termination check every N windows  */
	double bucket_width;		/* calendar bucket width in simulated time (power of two)          */
	uint32_t n_buckets;		/* calendar ring size (power of two)                               */
	uint32_t window_events;		/* target number of events per optimistic window                   */
	uint32_t arena_bytes;		/* per-LP state arena (model malloc heap), multiple of 16          */
	uint32_t max_payload;		/* largest event payload in bytes                                  */
	uint32_t flags;			/* RS_F_*                                                          */
	uint64_t max_pending;		/* capacity of the pending-event pool (events)                     */
	uint32_t max_window;		/* capacity of the per-window grouping buffer (events)             */
	uint32_t max_out;		/* capacity of the per-window output buffer (events)               */
	int32_t device;			/* CUDA device ordinal                                             */
	uint32_t topology_geometry;	/* filled from the model's topology_settings by rs_model_info()    */
	uint32_t topology_out_of;	/*   .out_of_topology                                              */
	uint32_t hot_threshold;		/* an LP with more window events (calendar + in-window arrivals) than this is a "chain LP":
					 * it gets a warp of its own in k_run_chain (0 = 256)                    */
	uint32_t run_ahead;		/* an LP whose window group starts with a burst of >= 256 same-timestamp events executes
					 * it at once, ahead of the pacing limit: 0/1 = yes (default), 2 = no   */
	uint32_t rank;			/* this engine's rank and the number of ranks (one engine per GPU);  */
	uint32_t nranks;		/*   nranks <= 1 means a single engine owning every LP              */
	uint32_t xchg_cap;		/* records per peer and exchange step (0 = 65536)                   */
	uint32_t chunk_events;		/* event budget of one visit of an LP: after this many events it stops and goes on
					 * in the next pacing round, while the other LPs move on (0 = 1536)      */
	uint32_t substeps;		/* pacing steps per window: the time limit below which LPs execute moves from the
					 * window start to its end in this many rounds (0 = chosen per window so that a
					 * step holds about one event per LP; 1 = no pacing inside a window)     */
	uint32_t pace_trail;		/* chain LPs stay one pacing step behind the others: 1 = yes, 0/2 = no (default; since
					 * arrivals are held back until the limit reaches them, trailing only costs rollbacks) */
	double window_span;		/* widest optimistic window in simulated time (0 = no limit besides
					 * window_events and 64 calendar buckets)                               */
	uint32_t burst_events;		/* budget of one visit for events that do not advance the LP's clock (a burst of
					 * same-timestamp events): after this many the visit stops, so that a burst does not
					 * set the length of the rounds it is executed in (0 = chunk_events)     */
	uint32_t helper_lanes;		/* the 31 idle lanes of a chain LP's warp copy, shift and scan its arrival array for it:
					 * 0/1 = yes (default), 2 = no                                            */
} rs_config_t;

/* Aggregate counters; definitions follow src/statistics/statistics.h:87-103 and :345-349. */
typedef struct rs_stats {
	uint64_t committed_events;	/* STAT_COMMITTED                                                  */
	uint64_t executed_events;	/* STAT_EVENT (every non-silent execution, rolled back ones too)   */
	uint64_t silent_events;		/* STAT_SILENT: coast-forward re-executions                        */
	uint64_t rollbacks;		/* STAT_ROLLBACK                                                   */
	uint64_t antimessages;		/* STAT_ANTIMESSAGE                                                */
	uint64_t checkpoints;		/* STAT_CKPT                                                       */
	uint64_t checkpoint_bytes;	/* STAT_CKPT_MEM                                                   */
	uint64_t windows;		/* GVT rounds                                                      */
	uint64_t relax_rounds;		/* in-window straggler rounds                                      */
	uint64_t late_events;		/* events delivered inside the window they were sent in           */
	uint64_t pending_events;	/* events still in the calendar                                    */
	uint64_t peak_window_events;
	double gvt;			/* last committed GVT (get_last_gvt(), src/gvt/gvt.c)              */
	double loop_seconds;		/* wall time of rs_dev_run loops (statistics_start..stop)          */
	double device_seconds;		/* CUDA-event time of rs_dev_run loops                             */
	uint32_t status;		/* RS_ST_* bits                                                    */
	uint32_t all_lps_agree;		/* last OnGVT AND-reduction (src/gvt/ccgs.c:64-84)                 */
	uint64_t kernel_launches;	/* engine kernels launched so far                                  */
	uint64_t chain_events;		/* events executed (silent ones included) by the chain-LP kernel k_run_chain */
	uint64_t sent_records;		/* records appended to the output buffer (events sent, antimessages included) */
	uint64_t reserved[4];		/* diagnostics: [0] groups sorted cooperatively, [1] LPs whose own-event queue moved to a heap,
					 * [2] run-ahead LPs (bursts executed ahead of the limit), [3] windows cut short by resource pressure */
} rs_stats_t;

/* Per-LP committed-trace digest (test/diagnostic interface; same definition as the reference-side
 * tracing shim in oracle/trace_shim.c): events with timestamp < simulation_time, in commit order. */
typedef struct rs_lp_trace {
	uint64_t count;			/* committed events of this LP                                     */
	uint64_t hash;			/* FNV-1a-64 over (timestamp bits, int32 type) of those events     */
	double last_ts;			/* timestamp of the last committed event                           */
	uint64_t seed;			/* RNG seed after it (numerical_state_t.seed)                      */
} rs_lp_trace_t;

typedef struct rs_model_info {
	const void *model_argp;		/* struct argp* of the model or NULL (src/ROOT-Sim.h:67-69)        */
	int32_t has_topology;		/* model defines topology_settings                                 */
	uint32_t topology_type;
	uint32_t topology_geometry;
	uint32_t topology_out_of;
	const char *topology_path;
	const char *engine;		/* "cuda-sm100a" (product) or "emu" (test-only build)              */
} rs_model_info_t;

typedef struct rs_engine rs_engine_t;

/* What the model linked into this library declares.  Replaces the weak-symbol probing of
 * src/core/init.c:385-390 (model_argp) and src/lib/topology/topology.c:210-214 (topology_settings). */
void rs_model_info(rs_model_info_t *out);

/* Hand model-specific command line options to the model's own argp parser (for callers that do not
 * go through host/rs_main.c).  argv[0] is skipped, like main()'s.  Returns 0 on success. */
int rs_model_args(int argc, char **argv);

/* Input files of the model.  The reference runs INIT on a CPU thread, so a model may fopen() a file there
 * (models/traffic/init.c:280-313 parses topology.txt once per LP).  Here INIT runs on the device: rs_dev_init
 * loads every file the model opens for reading under a literal name (found by rootsim-cc) from the working
 * directory -- where the reference's fopen would look; a missing file fails rs_dev_init with RS_ERR_MODEL, as
 * the model's own check would -- and the model's unchanged parsing code reads the image (fopen/fgets/rewind/
 * fclose/strtok_r/strtod ... of csrc/rs_prelude.cuh).  rs_model_file hands over the bytes of a file from HOST
 * memory instead (copied; data == NULL forgets the name); it must be called before rs_dev_init and takes
 * precedence over the working directory. */
int rs_model_file(const char *name, const void *data, size_t len);

/* Fill cfg with the defaults of the reference CLI (src/core/init.c:315-340) and of the engine. */
void rs_config_defaults(rs_config_t *cfg);

/* Per-LP seed derivation on the host: seeds[i] for global LP id lp_first+i.
 * Replaces numerical_init (src/lib/numerical.c:399-411) incl. sanitize_seed (:263-312). */
void rs_host_lp_seeds(uint64_t master_seed, uint32_t lp_first, uint32_t lp_count, uint64_t *seeds);

/* Create the device engine: allocate the LP table, state arenas, checkpoint store, calendar and
 * window buffers in HBM, copy the per-LP seeds (HOST buffer, lp_count entries; NULL = derive from
 * cfg->master_seed), and run INIT on every LP.  Replaces the parallel branch of SystemInit
 * (src/core/init.c:434-453) and initialize_worker_thread's INIT bootstrap (src/scheduler/scheduler.c:202-244).
 * Several engines of one library may be alive at a time (each call that runs model code first claims the
 * library's device-side engine descriptor for its engine); they must not be driven from different host threads
 * concurrently. */
int rs_dev_init(const rs_config_t *cfg, const uint64_t *host_seeds, rs_engine_t **out);

/* Advance the simulation by at most max_windows optimistic windows (0 = until termination).
 * One window = select -> group -> pacing rounds (wake, execute / roll back / coast-forward / cancel,
 * deliver) until nothing is left below the window end -> commit (GVT, fossil collection).  Replaces the body of main_simulation_loop
 * (src/main.c:132-170: process_bottom_halves, schedule, gvt_operations).
 * Returns RS_OK and sets *finished (may be NULL) when end_computing() (src/main.c:67-87) holds. */
int rs_dev_run(rs_engine_t *e, uint64_t max_windows, int *finished);

/* Counters (replaces statistics_stop's aggregation, src/statistics/statistics.c:442-568). */
int rs_dev_stats(rs_engine_t *e, rs_stats_t *out);

/* Copy n per-LP digests / committed-event counts / first `bytes` bytes of each LP's model state
 * to HOST buffers (per-LP statistics of src/statistics/statistics.c:660-733). */
int rs_dev_trace(rs_engine_t *e, rs_lp_trace_t *out, uint32_t n);
int rs_dev_lp_state(rs_engine_t *e, void *out, uint32_t bytes_per_lp, uint32_t n);

/*
 * Multi-GPU (one engine per GPU, LPs block-partitioned like src/core/core.c:275-300).  Every rank runs
 * rs_dev_run in lock step; per window the ranks agree on the window (two small all-reduces: GVT as a
 * MIN, bucket counts as a SUM -- the role of the MPI_Iallreduce(MIN) of src/communication/gvt.c:515-521)
 * and exchange, once per straggler round, the events/antimessages that land inside the window and,
 * once per window, the future events (role of the per-event MPI_Isend of src/communication/mpi.c:169-184).
 * The engine does not pick a transport: it calls the two collectives below on buffers in ENGINE
 * memory (device memory for the CUDA library).  rs_dev_comm_init_nccl installs an NCCL implementation
 * (ncclAllReduce + grouped ncclSend/ncclRecv on the engine's stream, NVLink/NVSwitch inside one box).
 */
typedef struct rs_comm_ops {
	void *ctx;
	/* in-place all-reduce of n int64; op 0 = sum, 1 = max */
	int (*allreduce_i64)(void *ctx, long long *buf, int n, int op, void *stream);
	/* personalised exchange: slot p of send goes to rank p, slot q of recv comes from rank q */
	int (*alltoall)(void *ctx, const void *send, void *recv, size_t bytes_per_peer, void *stream);
} rs_comm_ops_t;
int rs_dev_set_comm(rs_engine_t *e, const rs_comm_ops_t *ops);
/* id128: 128-byte ncclUniqueId created by rank 0 with rs_nccl_unique_id and handed to all ranks by the
 * launcher; libnccl_path NULL = "libnccl.so.2" (resolved with dlopen, no link-time dependency). */
int rs_nccl_unique_id(void *id128, const char *libnccl_path);
int rs_dev_comm_init_nccl(rs_engine_t *e, const void *id128, const char *libnccl_path);
/* A communicator shared by successive engines of one process (creation is collective and slow). */
int rs_nccl_comm_create(const void *id128, int rank, int nranks, const char *libnccl_path, void **comm_out);
int rs_dev_set_nccl_comm(rs_engine_t *e, void *comm);
int rs_nccl_comm_destroy(void *comm);
/* Direct exchange over NVLink / NVSwitch peer memory (ranks = processes on ONE box): after rs_dev_init and
 * the communicator, every rank exports a 64-byte handle of its inbox, the launcher gathers the nranks
 * handles (rank order) and hands the array to every rank.  From then on an exchange step is one kernel that
 * stores this rank's records straight into the peers' inboxes plus a one-word all-reduce as the barrier,
 * instead of the fixed-size personalised exchange of rs_comm_ops_t.alltoall (which remains the path for
 * other transports).  Engines of all ranks must be destroyed together (rs_dev_fini synchronises them). */
#define RS_IPC_HANDLE_BYTES 64
int rs_dev_ipc_export(rs_engine_t *e, void *handle64);
int rs_dev_ipc_import(rs_engine_t *e, const void *handles /* nranks * RS_IPC_HANDLE_BYTES; NULL = switch the direct exchange off again */);
/* this rank's block of LPs under the reference's block distribution (src/core/core.c:275-300) */
void rs_host_lp_block(uint32_t n_lps, uint32_t nranks, uint32_t rank, uint32_t *first, uint32_t *count);

/* Per-kernel device time (only with RS_F_KTIME): CUDA events recorded around every launch on the
 * engine's stream.  Returns the number of entries written. */
typedef struct rs_kernel_time {
	char name[32];
	double total_ms;
	uint64_t launches;
} rs_kernel_time_t;
int rs_dev_kernel_times(rs_engine_t *e, rs_kernel_time_t *out, int max_entries);

/* Diagnostic: the serial latency of the model's event handler on ONE device thread -- `iters` calls of
 * ProcessEvent(lp, now+k*1e-3, type, NULL, 0, state) back to back (silent != 0: as in a coast-forward, sends
 * suppressed).  This is what bounds a window whose longest per-LP chain dominates.  DESTRUCTIVE: the LP's state
 * advances; use a scratch engine.  ns_per_event receives the mean. */
int rs_dev_probe_event(rs_engine_t *e, uint32_t lp, int event_type, uint32_t iters, int silent, double *ns_per_event);

/* diagnostics on stderr: 1 = one line per straggler round that needed a host decision */
void rs_set_debug(int level);

const char *rs_dev_error(rs_engine_t *e);
int rs_dev_fini(rs_engine_t *e);

#ifdef __cplusplus
}
#endif
#endif
