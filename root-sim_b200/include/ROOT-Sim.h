/*
 * ROOT-Sim.h -- model-facing API of the B200-native Time Warp engine.
 *
 * This is the ONLY header a simulation model includes, exactly like the reference's
 * src/ROOT-Sim.h (contract: reference src/ROOT-Sim.h:52-222; man/ProcessEvent.3, man/OnGVT.3,
 * man/ScheduleNewEvent.3, man/SetState.3).  Model sources written against the reference header
 * compile against this one unchanged.  The same file is consumed in two ways:
 *
 *   RS_DEVICE_BUILD defined   (rootsim-cc -> nvcc, sm_100a): every API entry is a __device__
 *                             function implemented in csrc/rs_prelude.cuh; the model's
 *                             ProcessEvent/OnGVT become __device__ callbacks of the engine kernels.
 *   RS_DEVICE_BUILD undefined (plain C, gcc): ordinary external declarations.  This mode exists
 *                             for the CPU oracle under oracle/ (test infrastructure) only.
 *
 * Names, argument order/meaning and enum values are those of the reference API.
 */
#pragma once

#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include <stdbool.h>
#include <float.h>
#include <limits.h>
#include <argp.h>

#ifdef RS_DEVICE_BUILD
#define RS_API __host__ __device__
#else
#define RS_API
#endif

#ifdef __cplusplus
#define RS_WEAK_EXTERN extern
#else
#define RS_WEAK_EXTERN __attribute__((weak)) extern
#endif

/* Event code delivered to every LP at simulation start (reference ROOT-Sim.h:52-56). */
#ifdef INIT
#undef INIT
#endif
#define INIT 0

/* Simulation time (reference ROOT-Sim.h:58-59) and its +infinity (:61-62). */
typedef double simtime_t;
#define INFTY DBL_MAX

/* Total number of LPs in the run (reference ROOT-Sim.h:64-65). */
#ifdef RS_DEVICE_BUILD
#define n_prc_tot (rs_n_prc_tot())
RS_API unsigned int rs_n_prc_tot(void);
#else
extern unsigned int n_prc_tot;
#endif

/* Optional model command line parser, chained as an argp child (reference ROOT-Sim.h:67-69). */
RS_WEAK_EXTERN struct argp model_argp;

/* ---- per-LP rollbackable random numbers (reference ROOT-Sim.h:71-79, src/lib/numerical.c) ---- */
RS_API double Random(void);
RS_API int RandomRange(int min, int max);
RS_API int RandomRangeNonUniform(int x, int min, int max);
RS_API double Expent(double mean);
RS_API double Normal(void);
RS_API double Gamma(int ia);
RS_API double Poisson(void);
RS_API int Zipf(double skew, int limit);

/* ---- core API (reference ROOT-Sim.h:81-83) ---- */
#ifdef RS_DEVICE_BUILD
/* In the reference this is a function-pointer global switched between the serial and the parallel
 * engine (src/core/init.c:425,434); on the device there is one engine, so it is a plain function. */
RS_API void ScheduleNewEvent(unsigned int receiver, simtime_t timestamp, unsigned int event_type,
			     void *event_content, unsigned int event_size);
#else
extern void (*ScheduleNewEvent)(unsigned int receiver, simtime_t timestamp, unsigned int event_type,
				void *event_content, unsigned int event_size);
#endif
RS_API void SetState(void *new_state);

/* ---- topology library (reference ROOT-Sim.h:85-187) ---- */
enum _topology_geometry_t {
	TOPOLOGY_GEOMETRY_OFFSET = 1000,
	TOPOLOGY_HEXAGON = TOPOLOGY_GEOMETRY_OFFSET,
	TOPOLOGY_SQUARE,
	TOPOLOGY_RING,
	TOPOLOGY_BIDRING,
	TOPOLOGY_TORUS,
	TOPOLOGY_STAR,
	TOPOLOGY_GRAPH,
};

typedef enum _direction_t {
	DIRECTION_E = 0,
	DIRECTION_W = 1,
	DIRECTION_N = 2,
	DIRECTION_S = 3,
	DIRECTION_NE = 2,
	DIRECTION_SW = 3,
	DIRECTION_NW = 4,
	DIRECTION_SE = 5,
	DIRECTION_INVALID = UINT_MAX
} direction_t;

enum _topology_type_t {
	TOPOLOGY_OBSTACLES,
	TOPOLOGY_COSTS,
	TOPOLOGY_PROBABILITIES
};

/* A model that defines `topology_settings` switches the topology library on. */
RS_WEAK_EXTERN struct _topology_settings_t {
	const char *const topology_path;
	const enum _topology_type_t type;
	const enum _topology_geometry_t default_geometry;
	const unsigned out_of_topology;
	const bool write_enabled;
} topology_settings;

RS_API double GetValueTopology(unsigned from, unsigned to);
RS_API void SetValueTopology(unsigned from, unsigned to, double value);
RS_API unsigned int FindReceiver(void);
RS_API unsigned int RegionsCount(void);
RS_API unsigned int DirectionsCount(void);
RS_API unsigned int NeighboursCount(unsigned int region);
RS_API unsigned int GetReceiver(unsigned int from, direction_t direction, bool reachable);
RS_API unsigned int FindReceiverToward(unsigned int to);
#ifndef __cplusplus
double ComputeMinTour(unsigned int source, unsigned int dest, unsigned int result[RegionsCount()]);
#else
RS_API double ComputeMinTour(unsigned int source, unsigned int dest, unsigned int *result);
#endif

/* ---- agent-based-model library (reference ROOT-Sim.h:189-222): declared for source
 * compatibility; no BASELINE configuration uses it and the engine does not implement it. ---- */
typedef unsigned long long agent_t;

RS_WEAK_EXTERN struct _abm_settings_t {
	const unsigned neighbour_data_size;
	const unsigned traverse_handler;
	const bool keep_history;
} abm_settings;

RS_API int GetNeighbourInfo(direction_t i, unsigned int *region_id, void **data_p);
RS_API void TrackNeighbourInfo(void *neighbour_data);
RS_API bool IterAgents(agent_t *agent_p);
RS_API unsigned CountAgents(void);
RS_API agent_t SpawnAgent(unsigned user_data_size);
RS_API void KillAgent(agent_t agent);
RS_API void *DataAgent(agent_t agent, unsigned *data_size_p);
RS_API void ScheduleNewLeaveEvent(simtime_t time, unsigned int event_type, agent_t agent);
RS_API unsigned CountVisits(const agent_t agent);
RS_API void GetVisit(const agent_t agent, unsigned *region_p, unsigned *event_type_p, unsigned i);
RS_API void SetVisit(const agent_t agent, unsigned region, unsigned event_type, unsigned i);
RS_API void EnqueueVisit(agent_t agent, unsigned region, unsigned event_type);
RS_API void AddVisit(agent_t agent, unsigned region, unsigned event_type, unsigned i);
RS_API void RemoveVisit(agent_t agent, unsigned i);
RS_API unsigned CountPastVisits(const agent_t agent);
RS_API void GetPastVisit(const agent_t agent, unsigned *region_p, unsigned *event_type_p, simtime_t *time_p, unsigned i);
