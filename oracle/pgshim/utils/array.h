/* pgshim/utils/array.h -- intentionally minimal (see pgshim/postgres.h). */
#include "postgres.h"
#ifndef NDB_PGSHIM_ARRAY_H
#define NDB_PGSHIM_ARRAY_H
typedef struct ArrayType ArrayType;
#endif
