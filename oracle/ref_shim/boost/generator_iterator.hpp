// Stand-in for <boost/generator_iterator.hpp> (see oracle/ref_shim/README.md): included by the reference, nothing used.
#pragma once
