"""orbx: B200-native ORB front-end of Active-ORB-SLAM2 (host-side mirror of the reference's class surface)."""
