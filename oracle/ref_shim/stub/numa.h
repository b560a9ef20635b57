/* Stand-in for libnuma's <numa.h>. The reference includes it but never calls it
 * (SURVEY.md §2.1 row N). Test infrastructure only -- see oracle/README.md. */
#pragma once
