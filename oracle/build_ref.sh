#!/usr/bin/env bash
# oracle/build_ref.sh -- TEST INFRASTRUCTURE ONLY.
#
# Hand-builds the UNMODIFIED reference (HPDCS/ROOT-Sim 2.0.0) from the sources where they lie
# under $RS_REFERENCE (default /root/reference) into oracle/_ref/ (git-ignored; binaries only, no
# reference source is copied).  The reference's own build system (autotools) is NOT run; this is
# the recipe of SURVEY.md Appendix A:
#   * src/arch/asm_defines.h is generated with the rule of Makefile.am:34-36 into oracle/_ref/gen/
#     (found first through -I, the source tree stays untouched),
#   * librootsim / libdymelor / libwrapperl from the source lists of Makefile.am:44-84,136,138-143,
#     flags from configure.ac:23, no MPI,
#   * models linked exactly like scripts/rootsim-cc.in:189-265 (-D renames, ld -r, two --wrap passes).
# For every model two binaries are produced: <model> (plain) and <model>_trace (with
# oracle/trace_shim.c between runtime and model, dumping the per-LP trace digest).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
R="${RS_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$R/src" ]; then
	echo "build_ref.sh: reference tree $R not present; keeping prebuilt $OUT" >&2
	exit 0
fi
mkdir -p "$OUT/obj" "$OUT/gen/arch" "$OUT/lib"
printf '#define PACKAGE_STRING "ROOT-Sim 2.0.0"\n#define PACKAGE_BUGREPORT "rootsim@googlegroups.com"\n' > "$OUT/gen/cfg.h"
DEFS="-DOS_LINUX -DNDEBUG -DHAVE_LP_REBINDING -D_GNU_SOURCE -include $OUT/gen/cfg.h"
CF="-O3 -std=c11 -fno-omit-frame-pointer -ffloat-store -fno-common -fstrict-aliasing -fgnu89-inline -U_FORTIFY_SOURCE -pthread -w -I $OUT/gen -I $R/src"

# Makefile.am:34-36
gcc $CF $DEFS -DASM_DEFINES -S "$R/src/arch/asm_defines.c" -o - \
	| awk '($1=="->"){print "#define " $2 " " substr($3,2)}' > "$OUT/gen/arch/asm_defines.h"

ROOTSIM_SRC="main.c arch/memusage.c arch/thread.c arch/ult.c arch/x86.c arch/x86/disassemble.c arch/x86/jmp.S
 arch/x86/ecs_callback.S arch/x86/preempt_callback.S communication/wnd.c communication/communication.c
 communication/gvt.c communication/mpi.c core/init.c core/core.c datatypes/calqueue.c datatypes/hash_map.c
 datatypes/msgchannel.c gvt/gvt.c gvt/fossil.c gvt/ccgs.c lib/topology/topology.c lib/topology/costs.c
 lib/topology/obstacles.c lib/topology/probabilities.c lib/numerical.c lib/abm_layer.c lib/jsmn_helper.c
 lib/jsmn.c mm/state.c mm/ecs.c queues/queues.c queues/xxhash.c scheduler/binding.c scheduler/control.c
 scheduler/preempt.c scheduler/process.c scheduler/stf.c scheduler/scheduler.c serial/serial.c
 statistics/statistics.c"
DYMELOR_SRC="mm/checkpoints.c mm/platform.c mm/dymelor.c mm/buddy.c mm/segment.c mm/slab.c"
WRAPPER_SRC="lib-wrapper/wrapper.c"

build_lib() { # name, sources...
	local name="$1"; shift
	local objs=()
	for s in "$@"; do
		local o="$OUT/obj/${name}_$(echo "$s" | tr '/.' '__').o"
		if [ ! -f "$o" ] || [ "$R/src/$s" -nt "$o" ]; then
			gcc $CF $DEFS -c "$R/src/$s" -o "$o" &
		fi
		objs+=("$o")
	done
	wait
	rm -f "$OUT/lib/lib$name.a"
	ar cr "$OUT/lib/lib$name.a" "${objs[@]}"
}
build_lib rootsim $ROOTSIM_SRC
build_lib dymelor $DYMELOR_SRC
build_lib wrapperl $WRAPPER_SRC

build_model() { # model name, extra-cflags, suffix, extra objects
	local m="$1" extra="$2" sfx="$3"; shift 3
	local tmp; tmp="$(mktemp -d "$OUT/obj/link_${m}${sfx}.XXXX")"
	local objs=()
	for f in "$R"/models/"$m"/*.c; do
		local o="$tmp/$(basename "${f%.c}").o.rs"
		gcc $CF -c "$f" $extra -o "$o"
		objs+=("$o")
	done
	ld -r "${objs[@]}" "$@" -o "$tmp/APP.o"
	ld -r -L "$OUT/lib" --wrap strcpy --wrap strncpy --wrap strcat --wrap strncat --wrap memcpy --wrap memmove \
		--wrap memset --wrap strdup --wrap strndup -o "$tmp/APP-lib-wrapped.o" "$tmp/APP.o" -lwrapperl
	ld -r -L "$OUT/lib" --wrap malloc --wrap free --wrap realloc --wrap calloc -o "$tmp/APP-dym-wrapped.o" \
		"$tmp/APP-lib-wrapped.o" --whole-archive -ldymelor
	gcc "$tmp/APP-dym-wrapped.o" $CF -rdynamic -L "$OUT/lib" -lrootsim -lm -o "$OUT/${m}${sfx}" 2>/dev/null
	rm -rf "$tmp"
}

gcc $CF $DEFS -c "$HERE/trace_shim.c" -o "$OUT/obj/trace_shim.o"
for m in ${RS_REF_MODELS:-phold pcs traffic packet collector}; do
	build_model "$m" "-DOnGVT=OnGVT_light -DProcessEvent=ProcessEvent_light" ""
	build_model "$m" "-DOnGVT=OnGVT_model -DProcessEvent=ProcessEvent_model" "_trace" "$OUT/obj/trace_shim.o"
done
echo "build_ref.sh: built $(ls "$OUT" | grep -v -E '^(obj|gen|lib)$' | tr '\n' ' ')"
