/* pgshim/utils/memutils.h -- intentionally minimal (see pgshim/postgres.h). */
#include "postgres.h"
