"""Host-side mirror of the option-6 geometry block of pipeline/utils.py (full_prediction :517-574)."""
