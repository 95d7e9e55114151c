;;; Same public surface as the reference (package.lisp:26-49).
(defpackage "NAPA-FFT"
  (:use)
  (:export "COMPLEX-SAMPLE" "COMPLEX-SAMPLE-ARRAY"
           "REAL-SAMPLE" "REAL-SAMPLE-ARRAY"
           "DIRECTION" "SCALING" "WINDOWING"
           "%ENSURE-TWIDDLES"
           "%ENSURE-REVERSE" "GET-REVERSE"
           "%ENSURE-FFT" "GET-FFT" "GET-WINDOWED-FFT"
           "BIT-REVERSE"
           "FFT" "IFFT"
           "RECTANGULAR" "HANN" "BLACKMAN*"
           "BLACKMAN" "TRIANGLE" "BARTLETT"
           "GAUSS*" "GAUSSIAN" "GAUSSIAN*BARTLETT^X"
           "COSINE-SERIES" "BLACKMAN-HARRIS"
           "WINDOW-VECTOR"
           "WINDOWED-FFT" "WINDOWED-IFFT"
           "*RFFT-TWIDDLES*" "*RIFFT-TWIDDLES*"
           "%2RFFT"
           "RFFT" "RIFFT"
           "WINDOWED-RFFT" "WINDOWED-RIFFT"
           ;; new exports of the B200 build (batched / device-resident paths)
           "FFT-BATCH" "FILTER-CHUNKS" "NAPAFFT-ERROR"
           "DIST-INIT" "DIST-UNIQUE-ID" "DIST-FFT"))

(defpackage "NAPA-FFT.B200"
  (:use "CL" "SB-ALIEN" "SB-SYS" "NAPA-FFT"))
