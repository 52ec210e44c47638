/* pgshim/utils/jsonb.h -- intentionally minimal (see pgshim/postgres.h). */
#include "postgres.h"
#ifndef NDB_PGSHIM_JSONB_H
#define NDB_PGSHIM_JSONB_H
typedef struct Jsonb Jsonb;
typedef struct JsonbValue JsonbValue;
#endif
