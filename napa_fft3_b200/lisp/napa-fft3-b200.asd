;;; napa-fft3-b200.asd -- SBCL host layer of the B200 drop-in for Napa-FFT3.
;;; Same package name ("NAPA-FFT") and export list as the reference (package.lisp:26-49), so
;;; existing callers only change the system they load.  Requires libnapafft_b200.so
;;; (built by `python -m napa_fft3_b200.build`) on the loader path or in NAPAFFT_LIBRARY.
;;; NOTE: SBCL is not installed in the build image of this repository, so this layer is not
;;; exercised by the test-suite; napa_fft3_b200/host.py is the executed mirror of the same logic.
(asdf:defsystem "napa-fft3-b200"
  :version "0.1.0"
  :licence "BSD"
  :description "Napa-FFT3 easy interface on a B200 through the napafft C ABI"
  :serial t
  :components ((:file "package")
               (:file "ffi")
               (:file "easy-interface")
               (:file "windowing")
               (:file "real")))
