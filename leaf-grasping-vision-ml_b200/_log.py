"""Logging shim: the reference logs through rospy (out of scope); use it when present, else ``logging``."""
import logging

try:  # pragma: no cover - rospy is not installed in the build image
    import rospy as _rospy
    loginfo, logwarn, logerr = _rospy.loginfo, _rospy.logwarn, _rospy.logerr
except Exception:  # noqa: BLE001
    _log = logging.getLogger("leafgrasp_b200")
    loginfo, logwarn, logerr = _log.info, _log.warning, _log.error
