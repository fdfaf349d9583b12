/* colloc.h -- compatibility shim: the reference splits its public surface over
 * several headers (reference src/colloc.h); here everything lives in ntg.h. */
#ifndef NTG_DROPIN_COLLOC_SHIM_H_
#define NTG_DROPIN_COLLOC_SHIM_H_
#include "ntg.h"
#endif
