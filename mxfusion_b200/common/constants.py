"""Reserved prefixes shared by the inference layer (mirrors mxfusion/common/constants.py:15-16)."""
SET_PARAMETER_PREFIX = "SET_"
