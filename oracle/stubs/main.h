/* TEST INFRASTRUCTURE ONLY.  Empty stand-in for the STM32 board header that the
 * L432 copy of speechpy/feature.hpp includes (feature.hpp:32-33, "%%%SRH"); nothing
 * from it is used by the DSP code, so the host oracle build only needs it to exist. */
