// Forwarding header of the deal.II API shim (oracle/dealii_shim/deal.II/shim.h): NOT deal.II.
#include "../shim.h"
