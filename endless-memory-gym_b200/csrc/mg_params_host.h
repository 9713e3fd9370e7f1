// mg_params_host.h — host-only: option validation and the kernel parameter blocks derived from them.
#pragma once
#include "mg_mm_logic.h"
#include "mg_mp_logic.h"
#include "mg_ep_logic.h"
#include "mg_ss_logic.h"

void mg_set_err(const char* msg);
void mg_mm_derive(const mg_mm_params* p, MgMMK* k);
void mg_mp_derive(const mg_mp_params* p, MgMPK* k);
void mg_ss_derive(const mg_ss_params* p, MgSSK* k);
void mg_ep_derive(const mg_ep_params* p, MgEPK* k);
