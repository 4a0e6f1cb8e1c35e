// placeholder until the Barnes-Hut pipeline lands
#include "nb_internal.h"
namespace nb
{
int tree_reserve(nb_sim*) { set_error("Barnes-Hut mode is not built yet"); return NB_ERR_STATE; }
void tree_release(nb_sim*) {}
int tree_build(nb_sim*) { set_error("Barnes-Hut mode is not built yet"); return NB_ERR_STATE; }
int tree_walk(nb_sim*) { set_error("Barnes-Hut mode is not built yet"); return NB_ERR_STATE; }
}
extern "C" {
int nb_get_morton(nb_handle, uint64_t*, uint32_t*, size_t*) { nb::set_error("not built yet"); return NB_ERR_STATE; }
int nb_get_tree(nb_handle, int32_t*, int32_t*, int32_t*, double*, float*, size_t*) { nb::set_error("not built yet"); return NB_ERR_STATE; }
int nb_get_walk_stats(nb_handle, uint64_t*) { nb::set_error("not built yet"); return NB_ERR_STATE; }
}
