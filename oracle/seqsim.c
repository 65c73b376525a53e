/*
 * oracle/seqsim.c -- TEST INFRASTRUCTURE ONLY.  Never imported, linked or executed by the product
 * (root-sim_b200/, include/); used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.
 *
 * CPU restatement of the reference's SEQUENTIAL engine and of the library calls a model makes on
 * the hot path, small enough to run 1M+ LPs (the real reference cannot: MAX_LPs 250000, 4 MiB stack
 * per LP, O(#LP) lookup per event -- SURVEY.md top).  Parity is PINNED: tests/test_oracle.py
 * compares its per-LP trace digest bit-exactly with the unmodified reference built by
 * oracle/build_ref.sh (oracle/_ref/<model>_trace) and with the committed fixtures in tests/golden/.
 *
 * What is restated (reference file:line):
 *   event loop, INIT bootstrap, termination .... src/serial/serial.c:55-82 (SerialScheduleNewEvent),
 *                                                :84-105 (serial_init: INIT at ts 0 in gid order),
 *                                                :107-228 (loop, OnGVT bookkeeping, --simulation-time)
 *   ordering .................................... ascending timestamp, FIFO among equal timestamps
 *                                                (src/datatypes/calqueue.c:292-332) -- kept here by a
 *                                                binary heap on (timestamp, insertion counter)
 *   RNG ......................................... src/lib/numerical.c:54-72 (do_random), :80-134
 *                                                (Random, RandomRange, RandomRangeNonUniform, Expent),
 *                                                :141-249 (Normal, Gamma, Poisson, Zipf),
 *                                                :263-312 (sanitize_seed), :399-411 (per-LP seeds)
 *   topology .................................... src/lib/topology/obstacles.c:76-95 (free neighbour
 *                                                count), :171-213 (find_receiver_obstacles),
 *                                                src/lib/topology/topology.c:69-86 (directions),
 *                                                :183-203 (edge), :380-546 (get_raw_receiver)
 *
 * Build: gcc -O2 -ffp-contract=off -I root-sim_b200/include oracle/seqsim.c <model .c files compiled
 * with -DProcessEvent=ProcessEvent_light -DOnGVT=OnGVT_light> -lm     (see oracle/Makefile)
 * -DRS_ORACLE_PORTABLE_MATH selects rs_math.h's rs_log (bit-identical with the device build) instead
 * of libm's log (bit-identical with the reference on the same glibc).
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <string.h>
#include <stdbool.h>
#include <math.h>
#include <time.h>
#include <argp.h>

#include <ROOT-Sim.h>
#ifdef RS_ORACLE_PORTABLE_MATH
#include <rs_math.h>
#define ORACLE_LOG(x) rs_log(x)
#else
#define ORACLE_LOG(x) log(x)
#endif

extern void ProcessEvent_light(unsigned int me, simtime_t now, int event_type, void *event_content, unsigned int size, void *state);
extern bool OnGVT_light(unsigned int me, void *snapshot);

/* ------------------------------------------------------------------ types */
typedef struct event {
	double ts;
	double send_ts;
	uint64_t fifo;
	uint32_t sender, receiver;
	int32_t type;
	uint32_t size;
	struct event *next_free;
	unsigned char payload[];
} event_t;

typedef struct {
	uint64_t seed;		/* numerical.h:42-46 */
	double gset;
	bool iset;
	void *state;		/* SetState target */
	double lvt;
	/* trace digest (same definition as oracle/trace_shim.c) */
	uint64_t t_count, t_hash, t_seed;
	double t_last;
	unsigned char t_state[64];
	int t_have_state;
	/* topology */
	unsigned neighbours[6];
	unsigned free_neighbours;
} lp_t;

unsigned int n_prc_tot;
static void seq_schedule(unsigned int receiver, simtime_t timestamp, unsigned int event_type, void *event_content, unsigned int event_size);
void (*ScheduleNewEvent)(unsigned int, simtime_t, unsigned int, void *, unsigned int) = seq_schedule;

static lp_t *lps;
static lp_t *cur;
static unsigned cur_id;
static double cur_ts;

static double sim_time = 0;		/* --simulation-time, 0 = unbounded (init.c:167,336) */
static uint64_t master_seed = 0;
static bool have_seed = false;
static const char *output_dir = NULL;
static unsigned max_payload = 512;

/* ------------------------------------------------------------------ heap */
typedef struct { double ts; uint64_t fifo; event_t *e; } hnode;
static hnode *heap;
static size_t heap_n, heap_cap;
static uint64_t fifo_ctr;
static event_t *free_list;

static inline bool hless(const hnode *a, const hnode *b)
{
	return a->ts < b->ts || (a->ts == b->ts && a->fifo < b->fifo);
}

static void heap_push(event_t *e)
{
	if (heap_n == heap_cap) {
		heap_cap = heap_cap ? heap_cap * 2 : 1024;
		heap = realloc(heap, heap_cap * sizeof(hnode));
		if (!heap) { fprintf(stderr, "seqsim: out of memory\n"); exit(2); }
	}
	size_t i = heap_n++;
	hnode x = { e->ts, e->fifo, e };
	while (i > 0) {
		size_t p = (i - 1) >> 1;
		if (!hless(&x, &heap[p]))
			break;
		heap[i] = heap[p];
		i = p;
	}
	heap[i] = x;
}

static event_t *heap_pop(void)
{
	if (heap_n == 0)
		return NULL;
	event_t *top = heap[0].e;
	hnode x = heap[--heap_n];
	size_t i = 0;
	for (;;) {
		size_t c = 2 * i + 1;
		if (c >= heap_n)
			break;
		if (c + 1 < heap_n && hless(&heap[c + 1], &heap[c]))
			c++;
		if (!hless(&heap[c], &x))
			break;
		heap[i] = heap[c];
		i = c;
	}
	if (heap_n)
		heap[i] = x;
	return top;
}

static event_t *ev_alloc(unsigned size)
{
	if (size <= max_payload && free_list) {
		event_t *e = free_list;
		free_list = e->next_free;
		return e;
	}
	unsigned cap = size > max_payload ? size : max_payload;
	return malloc(sizeof(event_t) + cap);
}

static void ev_free(event_t *e)
{
	if (e->size > max_payload) { free(e); return; }
	e->next_free = free_list;
	free_list = e;
}

/* ------------------------------------------------------------------ core API */
/* serial.c:55-82 */
static void seq_schedule(unsigned int receiver, simtime_t timestamp, unsigned int event_type, void *event_content, unsigned int event_size)
{
	if (timestamp < cur_ts) {
		fprintf(stderr, "LP %u is trying to send events in the past. Current time: %f, scheduled time: %f\n", cur_id, cur_ts, timestamp);
		exit(EXIT_FAILURE);
	}
	if (receiver >= n_prc_tot) {
		/* the serial engine would crash in find_lp_by_gid; the parallel engine warns and drops
		 * (communication.c:280-283).  Follow the parallel engine. */
		fprintf(stderr, "Warning: ScheduleNewEvent to unknown LP %u dropped\n", receiver);
		return;
	}
	event_t *e = ev_alloc(event_size);
	e->ts = timestamp;
	e->send_ts = cur_ts;
	e->fifo = fifo_ctr++;
	e->sender = cur_id;
	e->receiver = receiver;
	e->type = (int32_t)event_type;
	e->size = event_size;
	if (event_size)
		memcpy(e->payload, event_content, event_size);
	heap_push(e);
}

void SetState(void *new_state)	/* state.c:321-324 */
{
	cur->state = new_state;
}

/* ------------------------------------------------------------------ RNG */
/* numerical.c:54-72 */
static double do_random(void)
{
	uint32_t s1 = (uint32_t)(cur->seed & 0xffffffffu);
	uint32_t s2 = (uint32_t)(cur->seed >> 32);
	s1 = 36969u * (s1 & 0xFFFFu) + (s1 >> 16);
	s2 = 18000u * (s2 & 0xFFFFu) + (s2 >> 16);
	cur->seed = ((uint64_t)s2 << 32) | s1;
	uint32_t r = (s1 << 16) + (s1 >> 16) + s2;
	return (r + 1.0) * 2.328306435454494e-10;
}

double Random(void) { return do_random(); }

int RandomRange(int min, int max)	/* numerical.c:91-100 */
{
	return (int)floor(do_random() * (max - min + 1)) + min;
}

int RandomRangeNonUniform(int x, int min, int max)	/* numerical.c:102-111 */
{
	int a = RandomRange(0, x);
	int b = RandomRange(min, max);
	return ((a | b) % (max - min + 1)) + min;
}

double Expent(double mean)	/* numerical.c:120-134 */
{
	if (mean < 0) {
		fprintf(stderr, "Expent() has been passed a negative mean value\n");
		exit(EXIT_FAILURE);
	}
	return -mean * ORACLE_LOG(1 - do_random());
}

double Normal(void)	/* numerical.c:141-169 */
{
	if (!cur->iset) {
		double v1, v2, rsq;
		do {
			v1 = 2.0 * Random() - 1.0;
			v2 = 2.0 * Random() - 1.0;
			rsq = v1 * v1 + v2 * v2;
		} while (rsq >= 1.0 || fabs(rsq) < DBL_EPSILON);	/* D_EQUAL_ZERO, core.h:94-96 */
		double fac = sqrt(-2.0 * ORACLE_LOG(rsq) / rsq);
		cur->gset = v1 * fac;
		cur->iset = true;
		return v2 * fac;
	}
	cur->iset = false;
	return cur->gset;
}

double Gamma(int ia)	/* numerical.c:179-216 */
{
	double x;
	if (ia < 1)
		ia = 1;
	if (ia < 6) {
		x = 1.0;
		for (int j = 1; j <= ia; j++)
			x *= Random();
		x = -ORACLE_LOG(x);
	} else {
		double v1, v2, y, am, s, e;
		do {
			do {
				do {
					v1 = Random();
					v2 = 2.0 * Random() - 1.0;
				} while (v1 * v1 + v2 * v2 > 1.0);
				y = v2 / v1;
				am = (double)(ia - 1);
				s = sqrt(2.0 * am + 1.0);
				x = s * y + am;
			} while (x < 0.0);
			e = (1.0 + y * y) * exp(am * ORACLE_LOG(x / am) - s * y);
		} while (Random() > e);
	}
	return x;
}

double Poisson(void) { return Gamma(1); }	/* numerical.c:223-226 */

int Zipf(double skew, int limit)	/* numerical.c:237-249 */
{
	double a = skew, b = pow(2., a - 1.), x, t, u, v;
	do {
		u = Random();
		v = Random();
		x = floor(pow(u, -1. / a - 1.));
		t = pow(1. + 1. / x, a - 1.);
	} while (v * x * (t - 1.) / (b - 1.) > (t / b) || x > limit);
	return (int)x;
}

/* numerical.c:263-312 */
static uint64_t sanitize_seed(uint64_t s)
{
	uint32_t lo = (uint32_t)s, hi = (uint32_t)(s >> 32), t;
	t = lo;
	if (t >= 0x9068FFFFu) t -= 0x9068FFFFu;
	if (t == 0) {
		t = lo ^ 0xFFFFFFFFu;
		if (t >= 0x9068FFFFu) t -= 0x9068FFFFu;
	}
	lo = t;
	t = hi;
	while (t >= 0x464FFFFFu) t -= 0x464FFFFFu;
	if (t == 0) {
		t = hi ^ 0xFFFFFFFFu;
		while (t >= 0x464FFFFFu) t -= 0x464FFFFFu;
	}
	hi = t;
	return ((uint64_t)hi << 32) | lo;
}

/* numerical.c:399-411: ROR on a SIGNED 64-bit value (arithmetic right shift); a shift count of 64
 * behaves as 0 on x86-64 (count masked to 6 bits), i.e. places == 0 yields the value itself. */
static uint64_t lp_seed(uint64_t master, unsigned gid)
{
	unsigned p = gid % 64;
	int64_t v = (int64_t)master;
	if (p == 0)
		return sanitize_seed((uint64_t)v);
	int64_t r = (int64_t)((uint64_t)v << p) | (v >> (64 - p));
	return sanitize_seed((uint64_t)r);
}

/* ------------------------------------------------------------------ topology */
static bool topo_on;
static unsigned topo_lp_cnt, topo_dirs, topo_edge;
static int topo_geom;

static unsigned raw_receiver(unsigned from, unsigned dir)	/* topology.c:380-546 */
{
	const unsigned INV = UINT_MAX;
	unsigned x, y, e = topo_edge, n = topo_lp_cnt;
	switch (topo_geom) {
	case TOPOLOGY_HEXAGON:
		y = from / e; x = from - y * e;
		switch (dir) {
		case DIRECTION_NW: x += (y & 1u) - 1; y -= 1; break;
		case DIRECTION_NE: x += (y & 1u); y -= 1; break;
		case DIRECTION_SW: x += (y & 1u) - 1; y += 1; break;
		case DIRECTION_SE: x += (y & 1u); y += 1; break;
		case DIRECTION_E: x += 1; break;
		case DIRECTION_W: x -= 1; break;
		default: return INV;
		}
		return (x < e && y < e) ? y * e + x : INV;
	case TOPOLOGY_SQUARE:
		y = from / e; x = from - y * e;
		switch (dir) {
		case DIRECTION_N: y -= 1; break;
		case DIRECTION_S: y += 1; break;
		case DIRECTION_E: x += 1; break;
		case DIRECTION_W: x -= 1; break;
		default: return INV;
		}
		return (x < e && y < e) ? y * e + x : INV;
	case TOPOLOGY_TORUS:
		y = from / e; x = from - y * e;
		switch (dir) {
		case DIRECTION_N: y -= 1; if (y >= e) y = e - 1; break;
		case DIRECTION_S: y += 1; if (y >= e) y = 0; break;
		case DIRECTION_E: x += 1; if (x >= e) x = 0; break;
		case DIRECTION_W: x -= 1; if (x >= e) x = e - 1; break;
		default: return INV;
		}
		return y * e + x;
	case TOPOLOGY_GRAPH:
		return dir < n ? dir : INV;
	case TOPOLOGY_BIDRING:
		if (dir == DIRECTION_W) return from == 0 ? n - 1 : from - 1;
		if (dir == DIRECTION_E) return from + 1 == n ? 0 : from + 1;
		return INV;
	case TOPOLOGY_RING:
		if (dir == DIRECTION_E) return from + 1 == n ? 0 : from + 1;
		return INV;
	case TOPOLOGY_STAR:
		if (from) return dir ? INV : 0;
		return dir + 1 >= n ? INV : dir + 1;
	}
	return INV;
}

static void topology_setup(void)
{
	topo_on = (&topology_settings != NULL);
	if (!topo_on)
		return;
	if (topology_settings.type != TOPOLOGY_OBSTACLES || topology_settings.topology_path) {
		fprintf(stderr, "seqsim: only TOPOLOGY_OBSTACLES without a topology file is restated\n");
		exit(2);
	}
	topo_lp_cnt = n_prc_tot - topology_settings.out_of_topology;
	topo_geom = topology_settings.default_geometry;
	switch (topo_geom) {	/* topology.c:69-86 */
	case TOPOLOGY_GRAPH: topo_dirs = topo_lp_cnt - 1; break;
	case TOPOLOGY_HEXAGON: topo_dirs = 6; break;
	case TOPOLOGY_TORUS: case TOPOLOGY_SQUARE: topo_dirs = 4; break;
	case TOPOLOGY_STAR: case TOPOLOGY_BIDRING: topo_dirs = 2; break;
	case TOPOLOGY_RING: topo_dirs = 1; break;
	default: topo_dirs = UINT_MAX;
	}
	topo_edge = 0;
	if (topo_geom == TOPOLOGY_SQUARE || topo_geom == TOPOLOGY_HEXAGON || topo_geom == TOPOLOGY_TORUS) {
		topo_edge = (unsigned)sqrt((double)topo_lp_cnt);	/* topology.c:183-203 */
		if (topo_edge * topo_edge != topo_lp_cnt) {
			fprintf(stderr, "Invalid number of regions for this topology geometry (must be a square number)\n");
			exit(EXIT_FAILURE);
		}
	}
	for (unsigned g = 0; g < n_prc_tot; g++) {	/* obstacles.c:76-95 */
		lp_t *l = &lps[g];
		unsigned i = topo_dirs;
		l->free_neighbours = i;
		if (topo_geom != TOPOLOGY_GRAPH) {
			while (i--) {
				unsigned id = raw_receiver(g, i);
				l->neighbours[i] = id;
				if (id == UINT_MAX)
					l->free_neighbours--;
			}
		} else if (topo_dirs && g < topo_dirs) {
			l->free_neighbours--;	/* the loop over i < directions meets i == gid once */
		}
	}
}

unsigned int FindReceiver(void)	/* topology.c:570-592 -> obstacles.c:171-213 */
{
	if (topo_lp_cnt <= cur_id) {
		fprintf(stderr, "FindReceiver(): source region %u when topology includes %u regions\n", cur_id, topo_lp_cnt);
		exit(EXIT_FAILURE);
	}
	if (!cur->free_neighbours)
		return cur_id;
	unsigned r;
	switch (topo_geom) {
	case TOPOLOGY_HEXAGON: case TOPOLOGY_SQUARE: case TOPOLOGY_TORUS: case TOPOLOGY_BIDRING: case TOPOLOGY_RING:
		do {
			r = cur->neighbours[(unsigned)(topo_dirs * Random())];
		} while (r == UINT_MAX);
		return r;
	case TOPOLOGY_GRAPH:
		return (unsigned)((topo_dirs + 1) * Random());
	case TOPOLOGY_STAR:
		if (cur_id != 0)
			return 0;
		return (unsigned)(((topo_lp_cnt - 1) * Random()) + 1);
	}
	return UINT_MAX;
}

unsigned int RegionsCount(void) { return topo_lp_cnt; }
unsigned int DirectionsCount(void) { return topo_dirs; }
unsigned int NeighboursCount(unsigned int region) { return lps[region].free_neighbours; }
unsigned int GetReceiver(unsigned int from, direction_t direction, bool reachable)
{
	(void)reachable;	/* no obstacles without a topology file */
	return raw_receiver(from, direction);
}

/* ------------------------------------------------------------------ trace digest */
static double horizon = 1e300;
static unsigned state_bytes = 8;
static uint64_t total_traced;

static inline uint64_t fnv_bytes(uint64_t h, const void *p, unsigned n)
{
	const unsigned char *c = p;
	for (unsigned i = 0; i < n; i++) {
		h ^= c[i];
		h *= 0x100000001b3ULL;
	}
	return h;
}

static void dump_trace(void)
{
	const char *path = getenv("RS_TRACE_OUT");
	if (!path)
		return;
	FILE *f = fopen(path, "w");
	if (!f) { perror(path); return; }
	fprintf(f, "# rs-trace v1 n_lp=%u horizon=%.17g state_bytes=%u total=%llu\n", n_prc_tot, horizon, state_bytes, (unsigned long long)total_traced);
	for (unsigned i = 0; i < n_prc_tot; i++) {
		lp_t *l = &lps[i];
		uint64_t tsb;
		memcpy(&tsb, &l->t_last, 8);
		fprintf(f, "%u %llu %016llx %016llx %016llx ", i, (unsigned long long)l->t_count, (unsigned long long)l->t_hash,
			(unsigned long long)tsb, (unsigned long long)l->t_seed);
		for (unsigned b = 0; b < state_bytes; b++)
			fprintf(f, "%02x", l->t_have_state ? l->t_state[b] : 0);
		fprintf(f, "\n");
	}
	fclose(f);
}

/* ------------------------------------------------------------------ CLI (subset of init.c:154-181) */
enum { OPT_LP = 1000, OPT_SIMT, OPT_SEED, OPT_DET, OPT_SEQ, OPT_OUT, OPT_WT, OPT_IGN1, OPT_IGN0 };
static const struct argp_option opts[] = {
	{"lp", OPT_LP, "N", 0, "Total number of LPs", 0},
	{"simulation-time", OPT_SIMT, "T", 0, "Stop at the first event with timestamp >= T", 0},
	{"seed", OPT_SEED, "S", 0, "Master seed (else $HOME/.rootsim/numerical.conf, else a constant)", 0},
	{"deterministic-seed", OPT_DET, 0, 0, "Accepted for CLI compatibility", 0},
	{"sequential", OPT_SEQ, 0, 0, "Accepted for CLI compatibility", 0},
	{"serial", OPT_SEQ, 0, OPTION_ALIAS, 0, 0},
	{"output-dir", OPT_OUT, "DIR", 0, "Write sequential_stats there", 0},
	{"wt", OPT_WT, "N", 0, "Accepted and ignored", 0},
	{0}
};

static error_t parse_opt(int key, char *arg, struct argp_state *st)
{
	(void)st;
	switch (key) {
	case OPT_LP: n_prc_tot = (unsigned)strtoul(arg, NULL, 10); break;
	case OPT_SIMT: sim_time = atof(arg); break;
	case OPT_SEED: master_seed = strtoull(arg, NULL, 10); have_seed = true; break;
	case OPT_OUT: output_dir = arg; break;
	case OPT_DET: case OPT_SEQ: case OPT_WT: break;
	default: return ARGP_ERR_UNKNOWN;
	}
	return 0;
}

int main(int argc, char **argv)
{
	struct argp_child kids[2] = { {0}, {0} };
	if (&model_argp)
		kids[0].argp = &model_argp;
	struct argp ap = { opts, parse_opt, NULL, "seqsim: CPU oracle (sequential engine restatement)", kids, NULL, NULL };
	argp_parse(&ap, argc, argv, 0, NULL, NULL);
	if (n_prc_tot == 0) {
		fprintf(stderr, "You must specify the total number of Logical Processes\n");
		return EXIT_FAILURE;
	}
	if (!have_seed) {	/* numerical.c:326-391: the master seed lives in ~/.rootsim/numerical.conf */
		char path[600];
		const char *home = getenv("HOME");
		master_seed = 12345678901234567ULL;
		if (home) {
			snprintf(path, sizeof path, "%s/.rootsim/numerical.conf", home);
			FILE *f = fopen(path, "r");
			if (f) {
				unsigned long long v;
				if (fscanf(f, "%llu", &v) == 1)
					master_seed = v;
				fclose(f);
			}
		}
	}
	const char *mp = getenv("RS_ORACLE_MAX_PAYLOAD");	/* event records are pooled at this payload size: 0 for
								 * payload-free models makes 16M-LP runs fit in memory */
	if (mp) max_payload = (unsigned)atoi(mp);
	const char *h = getenv("RS_TRACE_HORIZON");
	if (h) horizon = atof(h);
	const char *sb = getenv("RS_TRACE_STATE_BYTES");
	if (sb) state_bytes = (unsigned)atoi(sb);
	if (state_bytes > 64) state_bytes = 64;

	lps = calloc(n_prc_tot, sizeof(lp_t));
	if (!lps) { fprintf(stderr, "seqsim: out of memory\n"); return 2; }
	for (unsigned g = 0; g < n_prc_tot; g++) {
		lps[g].seed = lp_seed(master_seed, g);
		lps[g].t_hash = 0xcbf29ce484222325ULL;
	}
	topology_setup();

	/* serial.c:97-101: one INIT per LP at time 0, in gid order */
	for (unsigned g = 0; g < n_prc_tot; g++) {
		cur = &lps[g]; cur_id = g; cur_ts = 0.0;
		seq_schedule(g, 0.0, INIT, NULL, 0);
	}

	const long dbg_lp = getenv("RS_TRACE_LP") ? atol(getenv("RS_TRACE_LP")) : -1;
	bool *done = calloc(n_prc_tot, sizeof(bool));
	unsigned completed = 0;
	uint64_t executed = 0;
	bool finished = false;
	struct timespec t0, t1;
	clock_gettime(CLOCK_MONOTONIC, &t0);
	while (!finished) {
		event_t *e = heap_pop();
		if (!e) { fprintf(stderr, "No events to process!\n"); return EXIT_FAILURE; }
		cur_id = e->receiver; cur = &lps[cur_id]; cur_ts = e->ts;
		cur->lvt = e->ts;
		ProcessEvent_light(cur_id, e->ts, e->type, e->size ? e->payload : NULL, e->size, cur->state);
		executed++;
		if (dbg_lp == (long)cur_id && e->ts < horizon)	/* RS_TRACE_LP: list one LP's events (debugging aid) */
			fprintf(stderr, "[or lp %u] ts=%.17g\n", cur_id, e->ts);
		if (e->ts < horizon) {
			int32_t ty = e->type;
			cur->t_count++; total_traced++;
			cur->t_hash = fnv_bytes(cur->t_hash, &e->ts, 8);
			cur->t_hash = fnv_bytes(cur->t_hash, &ty, 4);
			cur->t_last = e->ts;
			cur->t_seed = cur->seed;
			if (cur->state) { memcpy(cur->t_state, cur->state, state_bytes); cur->t_have_state = 1; }
		}
		if (cur->state) {	/* serial.c:170-208, "normal" termination detection */
			bool d = OnGVT_light(cur_id, cur->state);
			if (done[cur_id] != d) { if (d) completed++; else completed--; }
			done[cur_id] = d;
			if (completed == n_prc_tot) finished = true;
		}
		if (sim_time > 0 && e->ts >= sim_time)	/* serial.c:211-213 */
			finished = true;
		ev_free(e);
	}
	clock_gettime(CLOCK_MONOTONIC, &t1);
	double wall = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
	printf("SEQSIM lps=%u executed=%llu wall_s=%.6f events_per_s=%.1f last_ts=%.17g\n", n_prc_tot,
	       (unsigned long long)executed, wall, executed / (wall > 0 ? wall : 1e-9), cur_ts);
	if (output_dir) {
		char p[700];
		snprintf(p, sizeof p, "mkdir -p '%s'", output_dir);
		if (system(p) == 0) {
			snprintf(p, sizeof p, "%s/sequential_stats", output_dir);
			FILE *f = fopen(p, "w");
			if (f) {
				fprintf(f, "TOTAL SIMULATION TIME ..... : %.3f seconds \nTOTAL LPs.................. : %u \nTOTAL EXECUTED EVENTS ..... : %llu \n",
					wall, n_prc_tot, (unsigned long long)executed);
				fclose(f);
			}
		}
	}
	dump_trace();
	return 0;
}
