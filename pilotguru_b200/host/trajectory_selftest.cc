// CPU test helper for trajectory.hpp (no GPU): trajectory_selftest <poses.json> <rotation_smooth_sigma> <out.json>
// poses.json = {"poses": [[time_usec, frame_id, tx, ty, tz, qw, qx, qy, qz], ...]}; runs the tail of TrackImageSequence.
#include "trajectory.hpp"

int main(int argc, char** argv) {
  PGB_CHECK(argc == 4) << "usage: trajectory_selftest poses.json sigma out.json";
  const std::string text = pgbhost::JsonReader::Slurp(argv[1]);
  pgbhost::JsonReader r(text);
  PGB_CHECK(r.FindTopLevel("poses"));
  std::vector<pgbhost::PoseWithTimestamp> tr;
  r.Expect('[');
  if (!r.TryConsume(']')) {
    do {
      r.Expect('[');
      pgbhost::PoseWithTimestamp p;
      p.time_usec = r.Int(); r.Expect(',');
      p.frame_id = r.Int(); r.Expect(',');
      double v[7];
      for (int i = 0; i < 7; i++) { v[i] = r.Double(); if (i < 6) r.Expect(','); }
      r.Expect(']');
      p.pose.t[0] = v[0]; p.pose.t[1] = v[1]; p.pose.t[2] = v[2];
      p.pose.qw = v[3]; p.pose.qx = v[4]; p.pose.qy = v[5]; p.pose.qz = v[6];
      p.is_lost = false;
      tr.push_back(p);
    } while (r.TryConsume(','));
    r.Expect(']');
  }
  return pgbhost::FinishTrajectory(tr, atoi(argv[2]), 0, argv[3], true) ? 0 : 3;
}
