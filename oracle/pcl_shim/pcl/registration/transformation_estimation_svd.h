// ORACLE ONLY: forwards to the Boost-free PCL restatement (see pcl/plade_pcl_shim.h).
#include <pcl/plade_pcl_shim.h>
