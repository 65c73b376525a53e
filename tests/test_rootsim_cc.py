"""The annotation pass of bin/rootsim-cc (file-scope functions -> RS_FN, variables -> RS_VAR)."""
import importlib.machinery
import importlib.util
import os

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_cc():
    path = os.path.join(REPO, "root-sim_b200", "bin", "rootsim-cc")
    loader = importlib.machinery.SourceFileLoader("rootsim_cc", path)
    spec = importlib.util.spec_from_loader("rootsim_cc", loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    return mod


SRC = r'''
#include <ROOT-Sim.h>
#define TWICE(x) ((x) * 2)   /* a ; in a comment; and a { brace */
typedef struct _s { int a; char *b; } s_t;
struct only_type { int q; };
enum color { RED, GREEN };
extern int shared_counter, other;
int shared_counter = 3, other = TWICE(2);
static double table[3] = { 1.0, 2.0, 3.0 };
const int K = 4;
char text[] = "semi;colon{";
void helper(s_t *p);
static inline int twice(int x) { return x * 2; }
struct argp model_argp = {0};
struct _topology_settings_t topology_settings = {.default_geometry = TOPOLOGY_GRAPH};
void (*callback)(int);
void ProcessEvent(unsigned int me, simtime_t now, int type, void *c, unsigned int size, void *state) {
	int local[4] = {1, 2, 3, 4};
	if (type == 1) { helper((s_t *)state); }
}
bool OnGVT(unsigned int me, void *snap) { return true; }
'''


def test_annotation():
    cc = load_cc()
    out, info = cc.annotate(SRC)
    assert info["functions"] == ["twice", "ProcessEvent", "OnGVT"]
    assert set(info["variables"]) >= {"shared_counter", "other", "table", "text"}
    assert "RS_VAR extern int shared_counter" in out
    assert "RS_VAR int shared_counter = 3" in out
    assert "RS_VAR static double table" in out
    assert "RS_FN void helper(s_t *p);" in out
    assert "RS_FN static inline int twice" in out
    assert "RS_VAR void (*callback)(int);" in out
    # host-only objects and types are left alone
    assert "RS_VAR struct argp" not in out and "RS_VAR struct _topology_settings_t" not in out
    assert "RS_VAR const int K" not in out
    assert "RS_VAR typedef" not in out and "RS_FN typedef" not in out
    assert "RS_VAR struct only_type" not in out and "RS_VAR enum" not in out
    # nothing inside a function body is touched
    assert "RS_VAR int local" not in out


def test_vla_rewrite():
    cc = load_cc()
    src = "void f(void) {\n\tbuffers *pointers[num_buffers];\n\tint fixed[8];\n}\n"
    out = cc.rewrite_vlas(src, {"num_buffers"})
    assert "rs_vla<buffers *, RS_VLA_CAP> pointers(num_buffers);" in out
    assert "int fixed[8];" in out
