/* pgshim/access/generic_xlog.h -- intentionally minimal (see pgshim/postgres.h). */
#include "postgres.h"
