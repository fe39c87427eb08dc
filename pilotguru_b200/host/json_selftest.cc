// Round-trip helper for the CPU tests of json_lite.hpp (no GPU, no libpgb200 calls):
//   json_selftest <in.json> <table> <value_field> <out.json> <root> <value_name>
// reads {table: [{value_field, time_usec, ...}]} and writes it back through JsonWriteTimestampedRealData.
#include "json_lite.hpp"

int main(int argc, char** argv) {
  PGB_CHECK(argc == 7) << "usage: json_selftest in table field out root value_name";
  const pgbhost::Table t = pgbhost::ReadTable(argv[1], argv[2], {argv[3]}, "time_usec");
  pgbhost::JsonWriteTimestampedRealData(t.integer, t.real[0], argv[4], argv[5], argv[6]);
  return 0;
}
