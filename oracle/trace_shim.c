/*
 * oracle/trace_shim.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Tracing shim linked between the UNMODIFIED reference runtime and an UNMODIFIED
 * reference model.  The model is compiled with
 *   -DProcessEvent=ProcessEvent_model -DOnGVT=OnGVT_model
 * and this file supplies the two symbols the reference runtime calls
 * (reference: src/core/core.h:221-222, stored per LP at src/scheduler/process.c:109-112).
 *
 * For every event with timestamp < horizon T (env RS_TRACE_HORIZON) it records, per LP,
 *   count, FNV-1a-64 over (timestamp bits, event type) in processing order,
 *   last timestamp, RNG seed after the event (current->numerical.seed,
 *   reference src/lib/numerical.h:42-46), first K bytes of the model state
 *   (env RS_TRACE_STATE_BYTES, default 8).
 * Only meaningful under --sequential.
 * The digest is written to the file named by env RS_TRACE_OUT at exit.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <stdbool.h>

#include <core/core.h>
#include <scheduler/process.h>
#include <scheduler/scheduler.h>
#include <mm/mm.h>

extern void ProcessEvent_model(unsigned int me, simtime_t now, int event_type, void *event_content, unsigned int size, void *state);
extern bool OnGVT_model(unsigned int me, void *snapshot);

#define RS_TRACE_MAX_STATE 64

struct lp_trace {
	uint64_t count;
	uint64_t hash;
	double last_ts;
	uint64_t seed;
	unsigned char state[RS_TRACE_MAX_STATE];
	int have_state;
};

static struct lp_trace *tr;
static unsigned tr_n;
static double horizon = 1e300;
static unsigned state_bytes = 8;
static uint64_t total;

static void dump(void)
{
	const char *path = getenv("RS_TRACE_OUT");
	if (!path || !tr)
		return;
	FILE *f = fopen(path, "w");
	if (!f)
		return;
	fprintf(f, "# rs-trace v1 n_lp=%u horizon=%.17g state_bytes=%u total=%llu\n", tr_n, horizon, state_bytes, (unsigned long long)total);
	for (unsigned i = 0; i < tr_n; i++) {
		uint64_t tsb;
		memcpy(&tsb, &tr[i].last_ts, 8);
		fprintf(f, "%u %llu %016llx %016llx %016llx ", i, (unsigned long long)tr[i].count,
			(unsigned long long)tr[i].hash, (unsigned long long)tsb, (unsigned long long)tr[i].seed);
		for (unsigned b = 0; b < state_bytes; b++)
			fprintf(f, "%02x", tr[i].have_state ? tr[i].state[b] : 0);
		fprintf(f, "\n");
	}
	fclose(f);
}

static void setup(void)
{
	tr_n = n_prc_tot;
	tr = rsalloc(tr_n * sizeof(*tr));	/* platform allocator: libc malloc is poisoned inside the runtime headers */
	memset(tr, 0, tr_n * sizeof(*tr));
	for (unsigned i = 0; i < tr_n; i++)
		tr[i].hash = 0xcbf29ce484222325ULL;
	const char *h = getenv("RS_TRACE_HORIZON");
	if (h)
		horizon = atof(h);
	const char *s = getenv("RS_TRACE_STATE_BYTES");
	if (s)
		state_bytes = (unsigned)atoi(s);
	if (state_bytes > RS_TRACE_MAX_STATE)
		state_bytes = RS_TRACE_MAX_STATE;
	atexit(dump);
}

static inline uint64_t fnv_bytes(uint64_t h, const void *p, unsigned n)
{
	const unsigned char *c = p;
	for (unsigned i = 0; i < n; i++) {
		h ^= c[i];
		h *= 0x100000001b3ULL;
	}
	return h;
}

void ProcessEvent_light(unsigned int me, simtime_t now, int event_type, void *event_content, unsigned int size, void *state)
{
	if (!tr)
		setup();
	ProcessEvent_model(me, now, event_type, event_content, size, state);
	if (now < horizon && me < tr_n) {
		struct lp_trace *t = &tr[me];
		int32_t ty = event_type;
		t->count++;
		total++;
		t->hash = fnv_bytes(t->hash, &now, 8);
		t->hash = fnv_bytes(t->hash, &ty, 4);
		t->last_ts = now;
		t->seed = current->numerical.seed;
		/* the state pointer may have been (re)set by this very event (INIT) */
		void *st = current->current_base_pointer;
		if (st) {
			memcpy(t->state, st, state_bytes);
			t->have_state = 1;
		}
	}
}

bool OnGVT_light(unsigned int me, void *snapshot)
{
	return OnGVT_model(me, snapshot);
}
