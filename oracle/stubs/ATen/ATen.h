// empty stand-in: the reference .cu includes this header but uses nothing from it (raw-pointer launchers only)
