// mg_ss_split.h — host entry of the two-kernel Searing Spotlights path (mg_ss_split.cu).
#pragma once
#include <stddef.h>

#include "mg_policy.h"

#define MG_SS_DESC_BYTES 256   // one frame description per env-step
#define MG_SS_SPLIT_CHUNK 8    // steps per logic / raster launch pair
#define MG_SS_SPLIT_TW 3       // warps per raster team (default)
#define MG_SS_RASTER_STAGE_MIN 18432  // shared memory kept for the staged asset prefix (default options: 18 080 B)

int mg_ss_split_launch(const MgSSK& k, const MgKIO& q, const void* assets, void* state, void* scratch, size_t scratch_bytes,
                       int tw_req, bool no_pdl, void* stream);
