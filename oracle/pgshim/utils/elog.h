/* pgshim/utils/elog.h -- intentionally minimal (see pgshim/postgres.h). */
#include "postgres.h"
