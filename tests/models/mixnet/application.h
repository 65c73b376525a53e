/* mixnet -- a small ROOT-Sim model written for this repository's tests (not derived from any
 * reference model).  It exercises what PHOLD does not: event payloads, zero-delay events between
 * neighbours, fixed-delay self events, model malloc/free inside events (rollbackable heap), and a
 * non-GRAPH topology (bidirectional ring). */
#include <ROOT-Sim.h>

enum { TICK = 1, TOKEN = 2, ZERO = 3, DROP = 4 };

typedef struct token {
	unsigned value;
	unsigned hops;
	simtime_t born;
} token_t;

typedef struct node {
	struct node *next;
	unsigned value;
} node_t;

typedef struct lp_state {
	unsigned ticks;
	unsigned tokens;
	unsigned zeros;
	unsigned n_nodes;
	unsigned long long sum;
	unsigned long long chain;
	node_t *head;
} lp_state_t;

extern double mean_delay;
extern double token_prob;
extern int order_sensitive;
extern unsigned tick_limit;

void push_node(lp_state_t *s, unsigned value);
void pop_node(lp_state_t *s);
