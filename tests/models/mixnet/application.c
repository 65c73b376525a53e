#include <math.h>
#include "application.h"

double mean_delay = 4.0;
double token_prob = 0.5;
int order_sensitive = 0;
unsigned tick_limit = 1000;

enum { OPT_MEAN = 300, OPT_PROB, OPT_ORDER, OPT_LIMIT };

const struct argp_option model_options[] = {
	{"mean", OPT_MEAN, "DOUBLE", 0, "mean TICK inter-arrival time", 0},
	{"token-prob", OPT_PROB, "DOUBLE", 0, "probability that a TICK emits a TOKEN", 0},
	{"order-sensitive", OPT_ORDER, 0, 0, "make TOKEN handling depend on the arrival order", 0},
	{"tick-limit", OPT_LIMIT, "UINT", 0, "OnGVT is satisfied after this many TICKs", 0},
	{0}
};

static error_t model_parse(int key, char *arg, struct argp_state *state)
{
	(void)state;
	switch (key) {
	case OPT_MEAN: mean_delay = atof(arg); break;
	case OPT_PROB: token_prob = atof(arg); break;
	case OPT_ORDER: order_sensitive = 1; break;
	case OPT_LIMIT: tick_limit = (unsigned)atoi(arg); break;
	default: return ARGP_ERR_UNKNOWN;
	}
	return 0;
}

struct argp model_argp = {model_options, model_parse, NULL, NULL, NULL, NULL, NULL};
struct _topology_settings_t topology_settings = {.default_geometry = TOPOLOGY_BIDRING};

void push_node(lp_state_t *s, unsigned value)
{
	node_t *n = malloc(sizeof(node_t));
	n->value = value;
	n->next = s->head;
	s->head = n;
	s->n_nodes++;
}

void pop_node(lp_state_t *s)
{
	node_t *n = s->head;
	if (!n)
		return;
	s->head = n->next;
	s->n_nodes--;
	free(n);
}

void ProcessEvent(unsigned int me, simtime_t now, int event_type, token_t *content, unsigned int size, lp_state_t *s)
{
	token_t tok;
	double u;
	simtime_t when;
	unsigned to;
	(void)size;
	/* NB: every RNG draw is sequenced through a local variable.  Drawing inside an argument list
	 * (f(FindReceiver(), now + Expent(x))) has unspecified evaluation order in C: gcc and nvcc
	 * order the two draws differently, and the trace would depend on the compiler. */

	switch (event_type) {
	case INIT:
		s = malloc(sizeof(lp_state_t));
		memset(s, 0, sizeof(lp_state_t));
		SetState(s);
		when = 10 * Random();
		ScheduleNewEvent(me, when, TICK, NULL, 0);
		break;

	case TICK:
		s->ticks++;
		when = now + Expent(mean_delay);
		ScheduleNewEvent(me, when, TICK, NULL, 0);
		u = Random();
		if (u < token_prob) {
			tok.value = (unsigned)(1000000 * Random());
			tok.hops = 0;
			tok.born = now;
			to = FindReceiver();
			when = now + Expent(mean_delay / 2);
			ScheduleNewEvent(to, when, TOKEN, &tok, sizeof(tok));
		}
		if (u > 0.9) {
			to = FindReceiver();
			ScheduleNewEvent(to, now, ZERO, NULL, 0);
		}
		push_node(s, s->ticks);
		if (s->n_nodes > 8)
			pop_node(s);
		break;

	case TOKEN:
		s->tokens++;
		s->sum += content->value;
		if (order_sensitive)
			s->chain = s->chain * 1000003ULL + content->value;
		else
			s->chain += (unsigned long long)content->value * content->value;
		if (content->hops < 3) {
			tok.value = order_sensitive ? content->value + 1 : (unsigned)(1000000 * Random());
			tok.hops = content->hops + 1;
			tok.born = content->born;
			to = FindReceiver();
			when = now + Expent(mean_delay / 4);
			ScheduleNewEvent(to, when, TOKEN, &tok, sizeof(tok));
		}
		break;

	case ZERO:
		s->zeros++;
		if (Random() < 0.5)
			ScheduleNewEvent(me, now + 1.0, DROP, NULL, 0);
		break;

	case DROP:
		pop_node(s);
		break;

	default:
		printf("mixnet: unknown event type %d\n", event_type);
		abort();
	}
}

bool OnGVT(unsigned int me, lp_state_t *snapshot)
{
	(void)me;
	return snapshot->ticks >= tick_limit;
}
