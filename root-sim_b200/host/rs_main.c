/*
 * rs_main.c -- host runtime of the B200-native engine: owns main(), parses the ROOT-Sim command
 * line, drives the device engine through the C-ABI of include/rootsim_b200.h and writes the
 * statistics file.  Plain C; no CUDA here (the reference's runtime is C too and owns main():
 * src/main.c:190-230).
 *
 * CLI: same option names and meaning as the reference (src/core/init.c:154-181, defaults :315-340).
 * Options that configure CPU-only machinery are accepted and ignored so that existing command lines
 * keep working (--wt, --no-core-binding, --scheduler, --lps-distribution, --gvt).  The model's own
 * options are chained as an argp child exactly like the reference does (src/core/init.c:385-390,419).
 * Extra `--b200-*` options size the device pools.
 *
 * Output: <output-dir>/execution_stats, thread_0_0/local_stats, thread_0_0/lps (--stats all|lp) with the labels of
 * src/statistics/statistics.c:345-394,483-505 and the
 * final "SIMULATION CORRECTLY COMPLETED" banner (:396-405); TIME BARRIER lines like src/main.c:150-165.
 */
#define _GNU_SOURCE
#include <argp.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdbool.h>
#include <time.h>
#include <sys/stat.h>

#include <rootsim_b200.h>

const char *argp_program_version = "ROOT-Sim B200 engine 0.1 (API of ROOT-Sim 2.0.0)";

enum {
	OPT_FIRST = 128, OPT_WT, OPT_LP, OPT_OUTPUT_DIR, OPT_SCHEDULER, OPT_NPWD, OPT_P, OPT_FULL, OPT_INC, OPT_A,
	OPT_GVT, OPT_CKTRM_MODE, OPT_GVT_SNAPSHOT_CYCLES, OPT_SIMULATION_TIME, OPT_LPS_DISTRIBUTION,
	OPT_DETERMINISTIC_SEED, OPT_SEED, OPT_SERIAL, OPT_SEQUENTIAL, OPT_NO_CORE_BINDING, OPT_VERBOSE, OPT_STATS,
	OPT_B_BUCKET, OPT_B_NBUCKETS, OPT_B_WINDOW, OPT_B_PENDING, OPT_B_ARENA, OPT_B_PAYLOAD, OPT_B_DEVICE, OPT_B_TRACE,
	OPT_B_SPAN, OPT_B_CHUNK, OPT_B_STEPS
};

static const struct argp_option options[] = {
	{"wt", OPT_WT, "VALUE", 0, "Number of worker threads (accepted; the device runs all LPs)", 0},
	{"lp", OPT_LP, "VALUE", 0, "Total number of Logical Processes being launched at simulation startup", 0},
	{"output-dir", OPT_OUTPUT_DIR, "PATH", 0, "Path to a folder where execution statistics are stored", 0},
	{"scheduler", OPT_SCHEDULER, "TYPE", 0, "LP Scheduling algorithm (accepted, ignored)", 0},
	{"npwd", OPT_NPWD, 0, 0, "Non Piece-Wise-Deterministic model: checkpoint before every event", 0},
	{"p", OPT_P, "VALUE", 0, "Checkpointing interval", 0},
	{"full", OPT_FULL, 0, 0, "Take only full logs", 0},
	{"inc", OPT_INC, 0, 0, "Take only incremental logs (not supported)", 0},
	{"A", OPT_A, 0, 0, "Autonomic subsystem (not supported)", 0},
	{"gvt", OPT_GVT, "VALUE", 0, "Time between two GVT reductions in ms (accepted; GVT advances every window)", 0},
	{"cktrm-mode", OPT_CKTRM_MODE, "TYPE", 0, "Termination Detection mode", 0},
	{"gvt-snapshot-cycles", OPT_GVT_SNAPSHOT_CYCLES, "VALUE", 0, "Termination detection is invoked after this number of windows", 0},
	{"simulation-time", OPT_SIMULATION_TIME, "VALUE", 0, "Halt the simulation when all LPs reach this logical time. 0 means infinite", 0},
	{"lps-distribution", OPT_LPS_DISTRIBUTION, "TYPE", 0, "LPs distribution over simulation kernels (block)", 0},
	{"deterministic-seed", OPT_DETERMINISTIC_SEED, 0, 0, "Do not change the master seed after the run", 0},
	{"seed", OPT_SEED, "VALUE", 0, "Manually specify the master seed", 0},
	{"serial", OPT_SERIAL, 0, 0, "Not available: this runtime is the optimistic device engine", 0},
	{"sequential", OPT_SEQUENTIAL, 0, 0, "Alias for --serial", 0},
	{"no-core-binding", OPT_NO_CORE_BINDING, 0, 0, "Accepted, ignored", 0},
	{"verbose", OPT_VERBOSE, "TYPE", 0, "Verbose execution (info, debug: a TIME BARRIER line per batch of windows; no: quiet), src/core/init.c:170", 0},
	{"stats", OPT_STATS, "TYPE", 0, "Level of detail in the output statistics", 0},
	{"b200-bucket-width", OPT_B_BUCKET, "T", 0, "Calendar bucket width in simulated time (power of two)", 1},
	{"b200-buckets", OPT_B_NBUCKETS, "N", 0, "Calendar ring size", 1},
	{"b200-window-events", OPT_B_WINDOW, "N", 0, "Target events per optimistic window", 1},
	{"b200-max-pending", OPT_B_PENDING, "N", 0, "Capacity of the pending-event pool", 1},
	{"b200-arena-bytes", OPT_B_ARENA, "N", 0, "Per-LP state arena in bytes", 1},
	{"b200-max-payload", OPT_B_PAYLOAD, "N", 0, "Largest event payload in bytes", 1},
	{"b200-device", OPT_B_DEVICE, "N", 0, "CUDA device ordinal", 1},
	{"b200-trace", OPT_B_TRACE, "FILE", 0, "Dump the per-LP committed-trace digest to FILE", 1},
	{"b200-window-span", OPT_B_SPAN, "T", 0, "Widest optimistic window in simulated time", 1},
	{"b200-chunk-events", OPT_B_CHUNK, "N", 0, "Event budget of one visit of an LP per pacing round", 1},
	{"b200-pace-steps", OPT_B_STEPS, "N", 0, "Pacing steps per window (0 = adaptive)", 1},
	{0}
};

static rs_config_t cfg;
static const char *output_dir = "outputs";
static const char *trace_file;
static bool deterministic_seed, seed_given, verbose;
static bool stats_lp = true;	/* --stats all (the reference's default) or lp: per-LP rows in thread_0_0/lps */

static int arena_given, payload_given;

/* The reference gives a model whatever it mallocs and sends (DyMeLoR grows, src/mm/dymelor.c:172-350; message
 * slabs of 512 bytes, src/communication/communication.c:431-449).  Here the per-LP heap and the payload slot are
 * fixed at engine creation.  When a run stops because one of them was too small and the user did not size it, the run is
 * started again from the beginning with twice the size (a simulation is a function of its seed: nothing is lost but
 * the time of the short run).  Returns 1 when a restart makes sense. */
static int grow_model_memory(unsigned status_bits)
{
	int again = 0;
	if ((status_bits & RS_ST_ARENA_FULL) && !arena_given && cfg.arena_bytes < (1u << 22)) {
		cfg.arena_bytes *= 2;
		fprintf(stderr, "[b200] the model's per-LP heap did not fit: restarting with --b200-arena-bytes %u\n", cfg.arena_bytes);
		again = 1;
	}
	if ((status_bits & RS_ST_PAYLOAD) && !payload_given && cfg.max_payload < 512u) {
		cfg.max_payload = cfg.max_payload ? 2 * cfg.max_payload : 16;
		fprintf(stderr, "[b200] an event payload did not fit: restarting with --b200-max-payload %u\n", cfg.max_payload);
		again = 1;
	}
	return again;
}

static error_t parse_opt(int key, char *arg, struct argp_state *state)
{
	switch (key) {
	case OPT_LP: cfg.n_lps = (uint32_t)strtoul(arg, NULL, 10); break;
	case OPT_OUTPUT_DIR: output_dir = arg; break;
	case OPT_P: {
		long v = strtol(arg, NULL, 10);
		if (v < 1 || v >= 40)	/* src/core/init.c:263-268 */
			argp_error(state, "Invalid value for the checkpointing interval (1..39)");
		cfg.ckpt_period = (uint32_t)v;
		break;
	}
	case OPT_NPWD: cfg.ckpt_period = 1; break;
	case OPT_INC: case OPT_A:
		argp_failure(state, EXIT_FAILURE, 38 /* ENOSYS, as src/core/init.c:273-279 */, "not supported");
		break;
	case OPT_GVT_SNAPSHOT_CYCLES: cfg.gvt_snapshot_cycles = (uint32_t)strtoul(arg, NULL, 10); break;
	case OPT_SIMULATION_TIME: cfg.simulation_time = (double)strtol(arg, NULL, 10); break;
	case OPT_DETERMINISTIC_SEED: deterministic_seed = true; break;
	case OPT_SEED: cfg.master_seed = strtoull(arg, NULL, 10); seed_given = true; break;
	case OPT_SERIAL: case OPT_SEQUENTIAL:
		argp_error(state, "--sequential is not available: this runtime is the optimistic B200 engine");
		break;
	case OPT_VERBOSE: verbose = strcmp(arg, "no") != 0; break;
	case OPT_STATS:		/* src/core/init.c:213-215: all | performance | lp | global */
		stats_lp = (!strcmp(arg, "all") || !strcmp(arg, "lp"));
		break;
	case OPT_WT: case OPT_SCHEDULER: case OPT_FULL: case OPT_GVT: case OPT_CKTRM_MODE: case OPT_LPS_DISTRIBUTION:
	case OPT_NO_CORE_BINDING:
		break;
	case OPT_B_BUCKET: cfg.bucket_width = atof(arg); break;
	case OPT_B_NBUCKETS: cfg.n_buckets = (uint32_t)strtoul(arg, NULL, 10); break;
	case OPT_B_WINDOW: cfg.window_events = (uint32_t)strtoul(arg, NULL, 10); break;
	case OPT_B_PENDING: cfg.max_pending = strtoull(arg, NULL, 10); break;
	case OPT_B_ARENA: cfg.arena_bytes = (uint32_t)strtoul(arg, NULL, 10); arena_given = 1; break;
	case OPT_B_PAYLOAD: cfg.max_payload = (uint32_t)strtoul(arg, NULL, 10); payload_given = 1; break;
	case OPT_B_DEVICE: cfg.device = atoi(arg); break;
	case OPT_B_TRACE: trace_file = arg; break;
	case OPT_B_SPAN: cfg.window_span = atof(arg); break;
	case OPT_B_CHUNK: cfg.chunk_events = (uint32_t)strtoul(arg, NULL, 10); break;
	case OPT_B_STEPS: cfg.substeps = (uint32_t)strtoul(arg, NULL, 10); break;
	case ARGP_KEY_END:
		if (cfg.n_lps == 0)
			argp_error(state, "Number of LPs was not provided \"--lp\"");
		break;
	default:
		return ARGP_ERR_UNKNOWN;
	}
	return 0;
}

/* Master seed file handling of src/lib/numerical.c:326-391: the seed lives in
 * $HOME/.rootsim/numerical.conf; unless --deterministic-seed is given it is replaced after loading. */
static uint64_t load_master_seed(void)
{
	char path[1024];
	const char *home = getenv("HOME");
	unsigned long long seed = 0;
	if (!home)
		home = ".";
	snprintf(path, sizeof path, "%s/.rootsim", home);
	mkdir(path, 0755);
	snprintf(path, sizeof path, "%s/.rootsim/numerical.conf", home);
	FILE *f = fopen(path, "r+");
	if (!f) {
		f = fopen(path, "w+");
		if (!f)
			return 0x5DEECE66DULL;
		FILE *r = fopen("/dev/urandom", "rb");
		if (r) {
			if (fread(&seed, sizeof seed, 1, r) != 1)
				seed = (unsigned long long)time(NULL);
			fclose(r);
		}
		seed &= 0x7fffffffffffffffULL;	/* seeds >= 2^63 collapse streams */
		fprintf(f, "%llu\n", seed);
		rewind(f);
	}
	if (fscanf(f, "%llu", &seed) != 1)
		seed = 0x5DEECE66DULL;
	if (!deterministic_seed) {
		rewind(f);
		srandom((unsigned)seed);
		fprintf(f, "%llu\n", (unsigned long long)random());
	}
	fclose(f);
	return seed;
}

static void stamp(char *buf, size_t n)
{
	time_t t = time(NULL);
	strftime(buf, n, "%Y-%m-%d %H:%M:%S", localtime(&t));
}

int main(int argc, char **argv)
{
	rs_model_info_t mi;
	rs_model_info(&mi);
	rs_config_defaults(&cfg);

	struct argp_child kids[2];
	memset(kids, 0, sizeof kids);
	if (mi.model_argp)
		kids[0].argp = (const struct argp *)mi.model_argp;
	struct argp ap = { options, parse_opt, NULL,
		"ROOT-Sim model running on the B200-native Time Warp engine", mi.model_argp ? kids : NULL, NULL, NULL };
	argp_parse(&ap, argc, argv, 0, NULL, NULL);

	if (!seed_given)
		cfg.master_seed = load_master_seed();
	if (!arena_given && cfg.n_lps) {
		/* default per-LP heap: 8 GiB of HBM shared by the LPs, between 256 bytes and 128 KiB each (a power of two);
		 * a checkpoint copies the used prefix only, so room that a model does not use costs memory, not bandwidth */
		unsigned long long a = (8ull << 30) / cfg.n_lps;
		uint32_t p = 256;
		while (p < (128u << 10) && 2ull * p <= a)
			p *= 2;
		cfg.arena_bytes = p;
	}

	char t_start[64], t_end[64];
	int attempts = 0;
restart:
	stamp(t_start, sizeof t_start);
	rs_engine_t *eng = NULL;
	int rc = rs_dev_init(&cfg, NULL, &eng);
	if (rc != RS_OK) {
		unsigned bits = 0;
		const char *msg = rs_dev_error(eng), *hex = msg ? strstr(msg, "status 0x") : NULL;
		if (hex)
			sscanf(hex + 7, "%x", &bits);
		if (attempts++ < 16 && grow_model_memory(bits)) {
			rs_dev_fini(eng);
			goto restart;
		}
		fprintf(stderr, "rs_dev_init failed (%d): %s\n", rc, rs_dev_error(eng));
		printf("\n--------- SIMULATION ABNORMALLY TERMINATED ----------\n");
		return EXIT_FAILURE;
	}
	printf("****************************\n*  ROOT-Sim Computing Node *\n****************************\n");
	printf("Engine: %s, LPs: %u, master seed: %llu, simulation time: %g\n", mi.engine, cfg.n_lps,
	       (unsigned long long)cfg.master_seed, cfg.simulation_time);

	int finished = 0;
	rs_stats_t st;
	double last_print = 0;
	/* rows of <output-dir>/thread_0_0/gvt (src/statistics/statistics.c:408-431): one per batch of windows */
	struct gvt_row { double wct, gvt; unsigned committed, cumulated; } *gvt_rows = NULL;
	size_t n_gvt_rows = 0, cap_gvt_rows = 0;
	unsigned long long prev_committed = 0;
	while (!finished) {
		rc = rs_dev_run(eng, 64, &finished);
		if (rc != RS_OK)
			break;
		rs_dev_stats(eng, &st);
		if (n_gvt_rows == cap_gvt_rows) {
			cap_gvt_rows = cap_gvt_rows ? 2 * cap_gvt_rows : 256;
			gvt_rows = realloc(gvt_rows, cap_gvt_rows * sizeof *gvt_rows);
		}
		if (gvt_rows) {
			gvt_rows[n_gvt_rows].wct = st.loop_seconds;
			gvt_rows[n_gvt_rows].gvt = st.gvt;
			gvt_rows[n_gvt_rows].committed = (unsigned)(st.committed_events - prev_committed);
			gvt_rows[n_gvt_rows].cumulated = (unsigned)st.committed_events;
			n_gvt_rows++;
			prev_committed = st.committed_events;
		}
		if (verbose || st.loop_seconds - last_print >= 1.0) {
			printf("TIME BARRIER %f - %llu committed events\n", st.gvt, (unsigned long long)st.committed_events);
			fflush(stdout);
			last_print = st.loop_seconds;
		}
	}
	rs_dev_stats(eng, &st);
	if (rc != RS_OK && attempts++ < 16 && grow_model_memory(st.status)) {
		free(gvt_rows);
		rs_dev_fini(eng);
		goto restart;
	}
	stamp(t_end, sizeof t_end);
	if (rc != RS_OK)
		fprintf(stderr, "simulation error (%d): %s\n", rc, rs_dev_error(eng));

	mkdir(output_dir, 0755);
	char path[1024];
	snprintf(path, sizeof path, "%s/thread_0_0", output_dir);
	mkdir(path, 0755);
	snprintf(path, sizeof path, "%s/thread_0_0/gvt", output_dir);
	FILE *fg = fopen(path, "w");
	if (fg) {	/* same header and row format as print_gvt_stats_file() */
		fprintf(fg, "#%15.15s    %15.15s    ", "\"WCT\"", "\"GVT VALUE\"");
		fprintf(fg, "%15.15s    %15.15s\n", "\"EVENTS\"", "\"CUMUL EVENTS\"");
		for (size_t i = 0; i < n_gvt_rows; i++) {
			fprintf(fg, " %15lf    %15lf    ", gvt_rows[i].wct, gvt_rows[i].gvt);
			fprintf(fg, "%15u    %15u\n", gvt_rows[i].committed, gvt_rows[i].cumulated);
		}
		fclose(fg);
	}
	free(gvt_rows);
	/* <output-dir>/execution_stats (node statistics) and <output-dir>/thread_0_0/local_stats (the statistics of the one
	 * "worker thread" there is: the device), labels of print_common_stats (src/statistics/statistics.c:345-394) */
	for (int which = 0; which < 2; which++) {
		snprintf(path, sizeof path, which ? "%s/thread_0_0/local_stats" : "%s/execution_stats", output_dir);
		FILE *f = fopen(path, "w");
		if (!f)
			continue;
		double tot = (double)st.executed_events + (double)st.silent_events;
		double rb_freq = tot > 0 ? (double)st.rollbacks / tot : 0;
		double rb_len = st.rollbacks ? (tot - (double)st.committed_events) / (double)st.rollbacks : 0;
		fprintf(f, "------------------------------------------------------------\n");
		fprintf(f, "-------------------- %s ---------------------\n", which ? "THREAD STATISTICS" : "GLOBAL STATISTICS");
		fprintf(f, "------------------------------------------------------------\n\n");
		if (!which) {
			fprintf(f, "SIMULATION STARTED AT ..... : %s \n", t_start);
			fprintf(f, "SIMULATION FINISHED AT .... : %s \n", t_end);
			fprintf(f, "TOTAL SIMULATION TIME ..... : %.03f seconds \n\n", st.loop_seconds);
		}
		fprintf(f, "TOTAL KERNELS ............. : %d \n", 1);
		if (which) {
			fprintf(f, "KERNEL ID ................. : %d \n", 0);
			fprintf(f, "LPs HOSTED BY KERNEL....... : %u \n", cfg.n_lps);
			fprintf(f, "TOTAL_THREADS ............. : %d \n", 1);
			fprintf(f, "THREAD ID ................. : %d \n", 0);
			fprintf(f, "LPs HOSTED BY THREAD ...... : %u \n\n", cfg.n_lps);
		} else {
			fprintf(f, "TOTAL_THREADS ............. : %d \n", 1);
			fprintf(f, "TOTAL LPs.................. : %u \n\n", cfg.n_lps);
		}
		fprintf(f, "TOTAL EXECUTED EVENTS ..... : %.0f \n", tot);
		fprintf(f, "TOTAL COMMITTED EVENTS..... : %llu \n", (unsigned long long)st.committed_events);
		fprintf(f, "TOTAL REPROCESSED EVENTS... : %llu \n", (unsigned long long)st.silent_events);
		fprintf(f, "TOTAL ROLLBACKS EXECUTED... : %llu \n", (unsigned long long)st.rollbacks);
		fprintf(f, "TOTAL ANTIMESSAGES......... : %llu \n", (unsigned long long)st.antimessages);
		fprintf(f, "ROLLBACK FREQUENCY......... : %.2f %%\n", rb_freq * 100);
		fprintf(f, "ROLLBACK LENGTH............ : %.2f events\n", rb_len);
		fprintf(f, "EFFICIENCY................. : %.2f %%\n", (1 - rb_freq * rb_len) * 100);
		fprintf(f, "AVERAGE EVENT COST......... : %.4f us\n", tot > 0 ? st.device_seconds * 1e6 / tot : 0);
		fprintf(f, "AVERAGE LOG SIZE........... : %.2f B\n", st.checkpoints ? (double)st.checkpoint_bytes / (double)st.checkpoints : 0);
		if (!which)
			fprintf(f, "\nLAST COMMITTED GVT ........ : %f\n", st.gvt);
		fprintf(f, "NUMBER OF GVT REDUCTIONS... : %llu\n", (unsigned long long)st.windows);
		fprintf(f, "SIMULATION TIME SPEED...... : %.2f units per GVT\n", st.windows ? st.gvt / (double)st.windows : 0);
		fprintf(f, "IN-WINDOW DELIVERIES....... : %llu\n", (unsigned long long)st.late_events);
		fprintf(f, "STRAGGLER ROUNDS........... : %llu\n", (unsigned long long)st.relax_rounds);
		fprintf(f, "DEVICE KERNEL LAUNCHES..... : %llu\n", (unsigned long long)st.kernel_launches);
		fprintf(f, "DEVICE STATUS BITS......... : 0x%x\n", st.status);
		fprintf(f, "\n--------- SIMULATION %s ----------\n", rc == RS_OK ? "CORRECTLY COMPLETED" : "ABNORMALLY TERMINATED");
		fclose(f);
	}
	if (stats_lp) {
		/* <output-dir>/thread_0_0/lps: one row per LP, header and columns of src/statistics/statistics.c:483-505.  The
		 * engine keeps the committed event count per LP; executed/rolled-back counts, rollbacks and antimessages exist as
		 * totals only (execution_stats), so those columns repeat the committed count or are 0. */
		rs_lp_trace_t *lt = malloc(sizeof(rs_lp_trace_t) * cfg.n_lps);
		snprintf(path, sizeof path, "%s/thread_0_0/lps", output_dir);
		FILE *fl = lt ? fopen(path, "w") : NULL;
		if (fl && rs_dev_trace(eng, lt, cfg.n_lps) == RS_OK) {
			fprintf(fl, "#%15.15s   %15.15s   %15.15s   %15.15s", "\"GID\"", "\"LID\"", "\"TOTAL EVENTS\"", "\"COMM EVENTS\"");
			fprintf(fl, "   %15.15s   %15.15s   %15.15s   %15.15s", "\"REPROC EVENTS\"", "\"ROLLBACKS\"", "\"ANTIMSG\"", "\"AVG EVT COST\"");
			fprintf(fl, "   %15.15s   %15.15s   %15.15s\n", "\"AVG CKPT COST\"", "\"AVG REC COST\"", "\"IDLE CYCLES\"");
			for (uint32_t i = 0; i < cfg.n_lps; i++)
				fprintf(fl, " %15u   %15u   %15.0lf   %15.0lf   %15.0lf   %15.0lf   %15.0lf   %15lf   %15.0lf   %15.0lf   %15.0lf   \n",
					i, i, (double)lt[i].count, (double)lt[i].count, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0);
		}
		if (fl)
			fclose(fl);
		free(lt);
	}
	if (trace_file) {
		rs_lp_trace_t *tr = malloc(sizeof(rs_lp_trace_t) * cfg.n_lps);
		unsigned char *stt = malloc(8u * cfg.n_lps);
		if (tr && stt && rs_dev_trace(eng, tr, cfg.n_lps) == RS_OK && rs_dev_lp_state(eng, stt, 8, cfg.n_lps) == RS_OK) {
			FILE *tf = fopen(trace_file, "w");
			if (tf) {
				fprintf(tf, "# rs-trace v1 n_lp=%u horizon=%.17g state_bytes=8 total=%llu\n", cfg.n_lps, cfg.simulation_time,
					(unsigned long long)st.committed_events);
				for (uint32_t i = 0; i < cfg.n_lps; i++) {
					unsigned long long tsb;
					memcpy(&tsb, &tr[i].last_ts, 8);
					fprintf(tf, "%u %llu %016llx %016llx %016llx ", i, (unsigned long long)tr[i].count,
						(unsigned long long)tr[i].hash, tsb, (unsigned long long)tr[i].seed);
					for (int b = 0; b < 8; b++)
						fprintf(tf, "%02x", stt[8u * i + b]);
					fprintf(tf, "\n");
				}
				fclose(tf);
			}
		}
		free(tr);
		free(stt);
	}
	rs_dev_fini(eng);
	printf("\n--------- SIMULATION %s ----------\n", rc == RS_OK ? "CORRECTLY COMPLETED" : "ABNORMALLY TERMINATED");
	return rc == RS_OK ? EXIT_SUCCESS : EXIT_FAILURE;
}
