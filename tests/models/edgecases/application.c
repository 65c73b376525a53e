/* edgecases -- a ROOT-Sim model written for this repository's tests: the error paths and degenerate inputs
 * of the event API.  --fault selects what one LP does wrong (or how little the model does at all). */
#include <ROOT-Sim.h>

enum { PING = 1 };
enum { F_NONE = 0, F_PAST = 1, F_RESERVED_TYPE = 2, F_BAD_RECEIVER = 3, F_EXIT = 4, F_ARENA = 5, F_SILENT_MODEL = 6, F_ONE_ACTIVE_LP = 7 };

int fault = F_NONE;
unsigned fault_lp = 3;

typedef struct { unsigned pings; unsigned sent; } state_t;

const struct argp_option model_options[] = {
	{"fault", 400, "INT", 0, "what to do wrong", 0},
	{"fault-lp", 401, "UINT", 0, "which LP does it", 0},
	{0}
};

static error_t model_parse(int key, char *arg, struct argp_state *state)
{
	(void)state;
	switch (key) {
	case 400: fault = atoi(arg); break;
	case 401: fault_lp = (unsigned)atoi(arg); break;
	default: return ARGP_ERR_UNKNOWN;
	}
	return 0;
}

struct argp model_argp = {model_options, model_parse, NULL, NULL, NULL, NULL, NULL};

void ProcessEvent(unsigned int me, simtime_t now, int event_type, void *content, unsigned int size, state_t *s)
{
	simtime_t when;
	(void)content; (void)size;
	if (event_type == INIT) {
		s = malloc(sizeof(state_t));
		s->pings = 0;
		s->sent = 0;
		SetState(s);
		if (fault == F_SILENT_MODEL)
			return;				/* a model that never schedules anything */
		if (fault == F_ONE_ACTIVE_LP && me != 0)
			return;				/* every LP but one stays idle for the whole run */
		when = 1.0 + Random();
		ScheduleNewEvent(me, when, PING, NULL, 0);
		return;
	}
	s->pings++;
	when = now + 0.5 + Random();
	if (me == fault_lp && s->pings == 5) {
		switch (fault) {
		case F_PAST: ScheduleNewEvent(me, now - 1.0, PING, NULL, 0); break;
		case F_RESERVED_TYPE: ScheduleNewEvent(me, when, 65533, NULL, 0); break;
		case F_BAD_RECEIVER: ScheduleNewEvent(n_prc_tot + 7, when, PING, NULL, 0); break;
		case F_EXIT: exit(3); break;
		case F_ARENA: while (malloc(64) != NULL) s->sent++; break;
		default: break;
		}
	}
	s->sent++;
	ScheduleNewEvent(fault == F_ONE_ACTIVE_LP ? me : (me + 1) % n_prc_tot, when, PING, NULL, 0);
}

bool OnGVT(unsigned int me, state_t *snapshot)
{
	(void)me;
	return snapshot->pings >= 1000000;
}
