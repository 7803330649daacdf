"""Constants of the path, same names and values as reference gypsum/constants.py:7-10,38 and gypsum/config.py:4-7."""
PRN_CHIP_COUNT = 1023  # constants.py:7
PRN_REPETITIONS_PER_SECOND = 1000  # constants.py:10
ONE_MILLISECOND = 0.001  # constants.py:38
ACQUISITION_INTEGRATION_PERIOD_MS = 10  # config.py:4
ACQUISITION_INTEGRATED_CORRELATION_STRENGTH_DETECTION_THRESHOLD = 3  # config.py:7
MILLISECONDS_TO_CONSIDER_FOR_TRACKER_LOCK_STATE = 250  # config.py:25
MAXIMUM_PHASE_ERROR_VARIANCE_FOR_LOCK_STATE = 900  # config.py:27
