;;; Keyword front-ends with the reference's signatures, defaults and aliasing rules
;;; (easy-interface.lisp:3-95, interface.lisp:81-217).  Only the arithmetic moved to the GPU.
(in-package "NAPA-FFT.B200")

(defun power-of-two-p (x) (= 1 (logcount x)))               ; support.lisp:61-62

(defun complex-samplify (vec &optional (size (length vec)))  ; support.lisp:64-93
  (etypecase vec
    (complex-sample-array vec)
    (sequence (map-into (make-array size :element-type 'complex-sample)
                        (lambda (x) (coerce x 'complex-sample)) vec))))
(defun real-samplify (vec &optional (size (length vec)))     ; support.lisp:95-112
  (etypecase vec
    (real-sample-array vec)
    (sequence (map-into (make-array size :element-type 'real-sample)
                        (lambda (x) (coerce x 'real-sample)) vec))))

;;; Aliasing contract of the easy interface (easy-interface.lisp:3-8, 44-57, 69-82 of the reference; it is
;;; part of the public behaviour and must not change): the transform runs in place on DST when DST is
;;; given (after copying the input there), in place on a FRESH vector when the caller's VEC had to be
;;; converted, and on a copy of VEC otherwise -- the caller's own complex-sample-array is never clobbered
;;; unless it is passed as :dst.
(defun copy-or-replace (src dst)
  (cond ((eql src dst) dst)
        (dst (replace dst src))
        (t (copy-seq src))))

(defun transform-target (vec dst n)
  "The vector a transform of the first N samples of VEC works on (see the aliasing contract above)."
  (let* ((samples (complex-samplify vec n))
         (target (if (or dst (eql samples vec))
                     (copy-or-replace samples dst)
                     samples)))             ; already a private conversion of the caller's sequence
    (assert (power-of-two-p n))
    (assert (>= (length samples) n))
    (assert (>= (length target) n))
    target))

(defun fft (vec &key dst size (in-order t) (scale nil) (window nil))
  (declare (type (or null complex-sample-array) dst))
  (let* ((n (or size (length vec)))
         (target (transform-target vec dst n)))
    (%c2c target target n 0 1 (scale-code scale) in-order window 0)))

(defun ifft (vec &key dst size (in-order t) (scale t) (window nil))
  (declare (type (or null complex-sample-array) dst))
  (let* ((n (or size (length vec)))
         (target (transform-target vec dst n)))
    (%c2c target target n 0 -1 (scale-code scale) in-order window 0)))

(defun bit-reverse (vec &optional dst (size (length vec)))
  (let ((elt (etypecase vec (complex-sample-array 0) (real-sample-array 1)))
        (dst (copy-or-replace vec dst)))
    (assert (>= (length vec) size))
    (assert (>= (length dst) size))
    (with-pinned-objects (dst)
      (check (napafft-bit-reverse (vector-sap dst) (vector-sap dst) size 1 size elt)))
    dst))

;;; low-level generators: closures over the C ABI instead of run-time compiled Lisp
(defun %ensure-twiddles (n forwardp)
  (assert (power-of-two-p n))
  (list :device-twiddles n forwardp))                        ; opaque: tables live on the device

(defun %ensure-fft (direction scaling windowing n)
  (check-type direction direction) (check-type scaling scaling) (check-type windowing windowing)
  (assert (power-of-two-p n))
  (let ((dir (direction-code direction)) (scale (scale-code scaling)))
    (if windowing
        (lambda (vec start window window-start twiddle)
          (declare (ignore twiddle))
          (%c2c vec vec n start dir scale nil window window-start))
        (lambda (vec start twiddle)
          (declare (ignore twiddle))
          (%c2c vec vec n start dir scale nil nil 0)))))

(defun %ensure-reverse (n &optional (eltype 'complex-sample))
  (assert (member eltype '(complex-sample real-sample)))
  (assert (power-of-two-p n))
  (let ((elt (if (eq eltype 'complex-sample) 0 1))
        (bytes (if (eq eltype 'complex-sample) 16 8)))
    (lambda (vec start)
      (with-pinned-objects (vec)
        (let ((sap (sb-sys:sap+ (vector-sap vec) (* bytes start))))
          (check (napafft-bit-reverse sap sap n 1 n elt))))
      vec)))

(defun get-reverse (n &optional (eltype 'complex-sample))
  (let ((fun (%ensure-reverse n eltype)))
    (lambda (vec) (funcall fun vec 0))))

(defun get-fft (size &key (forward t) (scale (if forward nil :inv)) (in-order t))
  (let ((dir (if forward 1 -1)) (scale (scale-code scale)))
    (lambda (vec) (%c2c vec vec size 0 dir scale in-order nil 0))))

(defun get-windowed-fft (size window-type &key (forward t) (scale (if forward nil :inv)) (in-order t))
  (assert window-type)
  (let ((dir (if forward 1 -1)) (scale (scale-code scale)))
    (lambda (vec window) (%c2c vec vec size 0 dir scale in-order window 0))))

;;; new export: B transforms of N samples stored back to back in one vector (H5 of SURVEY.md)
(defun fft-batch (vec n batch &key (direction 1) (in-order t) (scale nil) (window nil) (dst vec))
  (declare (type complex-sample-array vec dst))
  (assert (>= (length vec) (* n batch)))
  (with-pinned-objects (vec dst window)
    (check (napafft-c2c (vector-sap vec) (vector-sap dst) n batch n (direction-code direction)
                        (scale-code scale) (if in-order 1 0)
                        (if window (vector-sap window) (int-sap 0)) (window-code window) 0)))
  dst)


;;; example.lisp:81-114 -- the overlap-add streaming filter, every chunk in one device batch.
;;; Same arguments and result as the reference's FILTER-CHUNKS (a complex-sample-array of
;;; (length vector)); chunk energies are summed on the device, so the result agrees with the
;;; reference to rounding rather than bit for bit.
(defun filter-chunks (vector filter)
  (declare (type real-sample-array vector filter))
  (let* ((chunk-size  (length filter))
         (destination (make-array (length vector) :element-type 'complex-sample
                                                  :initial-element (complex 0d0)))
         (filter-rev  (bit-reverse filter))
         (window      (window-vector 'triangle chunk-size)))
    (sb-sys:with-pinned-objects (vector filter-rev window destination)
      (napa-fft.b200::check
       (napa-fft.b200::napafft-filter-chunks (sb-sys:vector-sap vector) (length vector)
                                              (sb-sys:vector-sap filter-rev) (sb-sys:vector-sap window)
                                              chunk-size (sb-sys:vector-sap destination))))
    destination))


;;; new export: ONE transform of 2^LGN points spread over the NRANKS processes that called DIST-INIT (BASELINE config 5).
;;; LOCAL-IN is this rank's block in the time layout ([N1][B] row-major, columns [rank*B, (rank+1)*B) of the [N1][N2] view
;;; of the vector) for :fwd, or in the frequency layout ([R][N2], rows k1 of X[k1 + N1*k2]) for :inv; the result is the other
;;; layout (include/napafft.h).  :inv scales by 1/N, like IFFT's default.
(defun dist-init (unique-id rank nranks)
  (declare (type (simple-array (unsigned-byte 8) (128)) unique-id))
  (with-pinned-objects (unique-id)
    (check (napafft-dist-init (vector-sap unique-id) rank nranks))))

(defun dist-unique-id ()
  (let ((id (make-array 128 :element-type '(unsigned-byte 8))))
    (with-pinned-objects (id)
      (check (napafft-dist-unique-id (vector-sap id))))
    id))

(defun dist-fft (local-in local-out lgn &key (direction 1))
  (declare (type complex-sample-array local-in local-out))
  (with-pinned-objects (local-in local-out)
    (check (napafft-dist-c2c (vector-sap local-in) (vector-sap local-out) lgn (direction-code direction))))
  local-out)
