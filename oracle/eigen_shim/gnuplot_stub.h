// Stand-in for tamcmc/headers/gnuplot-iostream.h (needs Boost.Iostreams; plotting is not on the hot path).
// Used with -DGNUPLOT_IOSTREAM_H -include gnuplot_stub.h so that tamcmc/headers/data.h compiles.
#ifndef ORACLE_GNUPLOT_STUB
#define ORACLE_GNUPLOT_STUB
#include <ostream>
namespace gnuplotio {
template <class T> struct TextSender { static void send(std::ostream& s, const T& v) { s << v; } };
}
#endif
