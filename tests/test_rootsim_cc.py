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


def test_cxx_keywords_used_as_c_identifiers_are_renamed():
    """models/traffic/functions.c:393 has a parameter called `new`; the unity source is C++."""
    cc = load_cc()
    src = 'void f(int id, double new) { /* new in a comment */ g(id, new); printf("new\\n"); }\nstruct s { int class; int newer; };\n'
    out = cc.rename_cxx_keywords(src)
    assert "double new_)" in out and "g(id, new_);" in out and "int class_;" in out
    assert "/* new in a comment */" in out and '"new\\n"' in out and "int newer;" in out


def test_input_files_of_a_model_are_found():
    """fopen(<literal or macro of one>, "r") names a file the engine loads before INIT (models/traffic/init.h:23,
    init.c:289); files opened for writing, or under a computed name, are not input images."""
    cc = load_cc()
    hdr = '#define FILENAME\t"topology.txt"\n#define OTHER 12\n'
    src = ('void init(void) { FILE *f = fopen(FILENAME, "r"); FILE *g = fopen("cfg/extra.dat", "rb");\n'
           ' FILE *w = fopen("out.txt", "w"); FILE *c = fopen(name, "r"); /* fopen("ghost", "r") */ }\n')
    assert cc.input_files([hdr, src]) == ["topology.txt", "cfg/extra.dat"]


def test_compile_only_objects_name_their_source(tmp_path, monkeypatch):
    """The reference's model Makefiles run `rootsim-cc -c x.c -o x.o` per file and link the objects
    (models/traffic/Makefile:13-17); here the link step rebuilds the unity source from what the objects name."""
    cc = load_cc()
    monkeypatch.chdir(tmp_path)
    (tmp_path / "a.c").write_text("int x;\n")
    (tmp_path / "b.c").write_text("int y;\n")
    assert cc.main(["-Wall", "-Wextra", "-c", "a.c", "-o", "a.o"]) == 0
    assert cc.main(["-c", "b.c"]) == 0
    assert (tmp_path / "a.o").read_text().split() == [cc.STUB_MAGIC, str(tmp_path / "a.c")]
    assert (tmp_path / "b.o").read_text().split() == [cc.STUB_MAGIC, str(tmp_path / "b.c")]
    (tmp_path / "foreign.o").write_bytes(b"\x7fELF....")
    assert cc.main(["-o", "model", "foreign.o"]) == 1


def test_unsupported_api_is_named():
    """ABM layer and costs/probabilities topology entries are declarations only: the wrapper says
    which one a model uses instead of leaving the user with a mangled name from the device linker."""
    cc = load_cc()
    src = "void f(void) { agent_t a = SpawnAgent(8); double v = GetValueTopology(0, 1); /* KillAgent(a) */ unsigned r = FindReceiver(); }\n"
    assert [nm for _, nm in cc.unsupported_api([src])] == ["GetValueTopology", "SpawnAgent"]
