// explicit instantiations, as in the reference (src/GOP.cpp:244-245)
#include "GOP.h"
template class GOPElement<float>;
template class GOPElement<double>;
template class GOP<float>;
template class GOP<double>;
