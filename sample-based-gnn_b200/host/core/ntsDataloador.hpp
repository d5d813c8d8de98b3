/* core/ntsDataloador.hpp -- drop-in for the reference header of the same name (AiX-im/Sample-based-GNN, core/ntsDataloador.hpp).
 *
 * Same mechanism as host/core/ntsFastSampler.hpp: the reference's own GNNDatum is included underneath, renamed, and `GNNDatum`
 * derives from it, so every member and method the toolkits use is the reference's own. What changes is
 *     readFeature_Label_Mask(feature_file, label_file, mask_file)          core/ntsDataloador.hpp:999-1063
 * which in the reference reads three text files token by token with operator>> on one thread (for cora, 3.9M floats: the
 * slowest part of start-up). Here the files are mapped and parsed by all host threads inside libnts_b200 (nb_read_feature_table /
 * nb_read_label_mask, csrc/loader.cu; strtof = the conversion operator>> ends in, so the values are bit-identical), and the parsed
 * table is cached next to the text file as raw floats (<feature_file>.nb_f32) for the following runs. NB_TEXT_LOADER=1 keeps the
 * reference's reader (A/B timing). Rows whose id is outside this partition are skipped, like the reference (:1030-1055).
 */
#ifndef NTS_B200_SHADOW_NTSDATALOADOR_HPP
#define NTS_B200_SHADOW_NTSDATALOADOR_HPP

#define GNNDatum NtsReferenceGNNDatum
#include_next "core/ntsDataloador.hpp"
#undef GNNDatum

#include <stdlib.h>

#include "nts_b200.h"

class GNNDatum : public NtsReferenceGNNDatum {
public:
  GNNDatum(GNNContext *_gnnctx, Graph<Empty> *graph_) : NtsReferenceGNNDatum(_gnnctx, graph_) {}

  void readFeature_Label_Mask(std::string inputF, std::string inputL, std::string inputM) {
    const char *legacy = getenv("NB_TEXT_LOADER");
    if (legacy && legacy[0] == '1') { NtsReferenceGNNDatum::readFeature_Label_Mask(inputF, inputL, inputM); return; }
    const double t0 = get_time();
    int from_cache = 0;
    static_assert(sizeof(long) == sizeof(int64_t), "labels are 64-bit");
    NTS_B200_CHECK(nb_read_feature_table(inputF.c_str(), graph->vertices, (uint32_t)gnnctx->layer_size[0], gnnctx->p_v_s, gnnctx->p_v_e,
                                         local_feature, 1, &from_cache));
    NTS_B200_CHECK(nb_read_label_mask(inputL.c_str(), inputM.c_str(), gnnctx->p_v_s, gnnctx->p_v_e, (int64_t *)local_label, (int32_t *)local_mask));
    printf("#feature/label/mask load: %.3f (s) [%s]\n", get_time() - t0, from_cache ? "binary cache" : "parallel text parse");
  }
};

#endif /* NTS_B200_SHADOW_NTSDATALOADOR_HPP */
